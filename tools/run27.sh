for g in 32 128; do
MP_L2_FETCH=$g timeout 300 python tools/mb_policies.py 0 && MP_L2_FETCH=$g timeout 500 ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum -k regex:k_gather_bench --csv --log-file gpurun_out/mb_ncu_$g.csv python tools/mb_policies.py 0 > gpurun_out/mb_$g.log 2>&1
done
