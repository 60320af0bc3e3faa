python - <<'PY'
import megapath_b200 as mp
c=mp.Context(0)
for k in (0,10,11,12,13,14,15,16,1,21,22,23,24):
    try: print(k, round(c.microbench(k),1), flush=True)
    except Exception as e: print(k,'err',e)
PY
