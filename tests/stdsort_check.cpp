// Checks megapath_b200/csrc/mp_stdsort.h against the real std::sort of the toolchain the reference is built with (libstdc++),
// element order of ties included.  Built and run by tests/test_stdsort.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../megapath_b200/csrc/mp_stdsort.h"

struct E { int key, id; };
struct Desc { bool operator()(const E &a, const E &b) const { return a.key > b.key; } };

static bool check(std::vector<E> v, const char *what)
{
    std::vector<E> a = v, b = v;
    std::sort(a.begin(), a.end(), Desc());
    mp_stdsort::sort(b.data(), b.data() + b.size(), Desc());
    for (size_t i = 0; i < v.size(); ++i)
        if (a[i].id != b[i].id) { printf("MISMATCH %s n=%zu at %zu: std (%d,%d) mine (%d,%d)\n", what, v.size(), i, a[i].key, a[i].id, b[i].key, b[i].id); return false; }
    return true;
}

int main()
{
    std::mt19937 rng(12345);
    long cases = 0;
    for (int n = 0; n <= 3000; n += (n < 70 ? 1 : n < 400 ? 7 : 131)) {
        for (int range : { 1, 2, 3, 5, 17, 100, 100000 }) {
            for (int rep = 0; rep < (n < 70 ? 40 : 6); ++rep) {
                std::vector<E> v(n);
                for (int i = 0; i < n; ++i) { v[i].key = (int)(rng() % (unsigned)range); v[i].id = i; }
                if (!check(v, "random")) return 1;
                ++cases;
            }
        }
        std::vector<E> v(n);
        for (int i = 0; i < n; ++i) { v[i].key = i; v[i].id = i; }
        if (!check(v, "ascending")) return 1;
        for (int i = 0; i < n; ++i) v[i].key = n - i;
        if (!check(v, "descending")) return 1;
        for (int i = 0; i < n; ++i) v[i].key = i < n / 2 ? i : n - i;
        if (!check(v, "organ pipe")) return 1;
        for (int i = 0; i < n; ++i) v[i].key = (i * 7919) % 13;
        if (!check(v, "few values")) return 1;
        cases += 4;
    }
    // median-of-three killer (Musser): drives introsort into its heapsort fallback
    for (int n : { 64, 256, 1024, 4096 }) {
        std::vector<E> v(n);
        const int k = n / 2;
        for (int i = 1; i <= k; ++i) {
            if (i % 2 == 1) { v[i - 1].key = -i; v[i].key = -(k + i); }
            v[k + i - 1].key = -(2 * i);
        }
        for (int i = 0; i < n; ++i) v[i].id = i;
        if (!check(v, "median-of-3 killer")) return 1;
        ++cases;
    }
    printf("ok %ld cases\n", cases);
    return 0;
}
