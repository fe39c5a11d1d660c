for l2 in 0 1 2 3; do echo -n "L2promo=$l2: "; BOWGPU_TMAP_L2=$l2 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_step %.4f  kernel_ms %.4f  GB/s %.0f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved']))"; done
ncu --set full --clock-control none --import-source on -k regex:segreduce_kernel -s 3 -c 1 -o gpurun_out/prof_seg_r1o python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
