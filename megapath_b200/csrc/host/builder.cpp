// builder.cpp -- `2bwt-builder <sequence file>` drop-in (runMegaPath.sh:303 indexes the assembled contigs with it): FASTA in,
// <fasta>.index.{pac,ann,amb,tra,bwt,fmv,sa,lkt} out, byte-identical to the reference's 2bwt-lib/2BWT-Builder.c with its shipped
// 2bwt-builder.ini (forward index only, OccValueFreq 256, SaValueFreq 16, 13-mer lookup table).  The FASTA front end is host code
// (fasta_index.h); suffix sorting, BWT, occurrence tables, SA samples and the lookup table are built in HBM (mp_index_build).
//   -U                  treat lower-case bases as ambiguous (MaskLowerCase, 2BWT-Builder.c:433)
//   --annotation-only   stop after .pac/.ann/.amb/.tra (no GPU needed)
#include <stdio.h>
#include <string.h>
#include "megapath_b200.h"
#include "fasta_index.h"

int main(int argc, char **argv)
{
    const char *fasta = nullptr; bool mask = false, annOnly = false; int device = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-U")) mask = true;
        else if (!strcmp(argv[i], "--annotation-only")) annOnly = true;
        else if (!strcmp(argv[i], "-c") && i + 1 < argc) device = atoi(argv[++i]);
        else if (argv[i][0] != '-' && !fasta) fasta = argv[i];
    }
    if (!fasta) { printf("Usage: ./%s <sequence file>\n", argv[0]); return 1; }
    const std::string prefix = std::string(fasta) + ".index";
    printf("Parsing FASTA file..\n");
    FastaText t; std::string err;
    if (!fasta_parse(fasta, mask, t)) { fprintf(stderr, "ParseFASTToPacked() : %s\n", t.error.c_str()); return 1; }
    if (!fasta_write_annotation(t, prefix, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    printf("Finished. Parsed %llu sequences.\n", (unsigned long long)t.seqs.size());
    if (annOnly) {
        FILE *f = fopen((prefix + ".pac").c_str(), "wb");
        if (!f) { fprintf(stderr, "cannot create %s.pac\n", prefix.c_str()); return 1; }
        fwrite(t.pac.data(), 1, t.pac.size(), f);
        if (t.n % 4 == 0) fputc(0, f);
        fputc((int)(t.n % 4), f);
        fclose(f);
        return 0;
    }
    if (t.n < 32) { fprintf(stderr, "the text has fewer than 32 bases\n"); return 1; }
    mp_context *ctx = nullptr;
    if (mp_init(device, &ctx)) { fprintf(stderr, "%s\n", mp_last_error()); return 1; }
    printf("Building BWT, occurrence tables, SA samples and the lookup table on the GPU..\n");
    if (mp_index_build(ctx, t.pac.data(), t.n) || mp_index_save(ctx, prefix.c_str())) { fprintf(stderr, "%s\n", mp_last_error()); mp_destroy(ctx); return 1; }
    mp_destroy(ctx);
    printf("Index building is completed.\n");
    return 0;
}
