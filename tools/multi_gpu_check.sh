timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_dropin.py tests/test_gpu_sfxiterator.py -x -q 2>&1 | tail -3
for N in 8 4 2; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/r2m_bench_c4_${N}gpu.json 2> gpurun_out/r2m_err_${N}gpu.txt
  python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench_c4_${N}gpu.json').read().strip().splitlines()[-1]); print($N, round(d['ms_per_step'],2), round(d['wall_ms_per_step'],2), d['breakdown_ms_last_step'], d['checks']['identical_to_reference'], round(d['e2e']['ms_per_step'],1), d['roofline']['frac'])"
  grep -v "OMP_NUM\|^\*\|^$" gpurun_out/r2m_err_${N}gpu.txt | tail -3
done
timeout 300 python tools/group_bench.py --gpus 8 --steps 4 --copy 2>/dev/null > gpurun_out/r2m_group8.json; python -c "
import json; d=json.loads(open('gpurun_out/r2m_group8.json').read()); print('group8', d['wall_ms_per_step'], d['job'], d['gather_ms'], d['identical_to_reference'])"
