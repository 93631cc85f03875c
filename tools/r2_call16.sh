mkdir -p gpurun_out
{
echo "== peer BN test (2 GPUs)"
timeout 300 python -m pytest tests/test_peer_bn_gpu.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error -e "^E " | tail -12
echo "== bench 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2b.json')); print({k: d.get(k) for k in ('value','ms_per_step','e2e','comm')}, d['config']['cuda_graph'])"
grep -v "^\[W\|^W1017\|warn\|Warn" gpurun_out/r2_bench_n2b.err | tail -8
} 2>&1 | tee gpurun_out/r2_call16.log
