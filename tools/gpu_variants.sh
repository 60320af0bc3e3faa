#!/bin/bash
# two builds of the K = 5 fill kernel (register budget for 6 or 5 blocks per SM), short cfg2 / cfg4 benches with each
for mb in 6 1; do
  (cd megapath_b200/csrc && touch mp_dp.cu && make -s EXTRA=-DMP_FILL_MINBLOCKS=$mb 2>&1 | tail -2; grep -A3 "k_dp_fillILi5ELin2" build/mp_dp.ptxas.log | grep Used)
  if [ $mb = 6 ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2; fi
  for cfg in cfg2 cfg4; do MP_BENCH_VERBOSE=1 timeout 300 python bench.py --config $cfg --no-cpu-baseline --steps 6 > gpurun_out/bv_${mb}_$cfg.json 2> gpurun_out/bv_${mb}_$cfg.err; echo "minblocks $mb $cfg: $(grep 'loop R' gpurun_out/bv_${mb}_$cfg.err | sed 's/.*ms_fill/ms_fill/')"; python -c "
import json; d=json.load(open('gpurun_out/bv_${mb}_$cfg.json')); print('   value', round(d['value']/1e6,2), 'gcups', round(d['roofline']['compute']['gcups_fill']))"; done
done
