mkdir -p gpurun_out
{
echo "== peer BN test (2 GPUs)"
timeout 300 python -m pytest tests/test_peer_bn_gpu.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error -e "^E " | tail -12
} 2>&1 | tee gpurun_out/r2_call16.log
