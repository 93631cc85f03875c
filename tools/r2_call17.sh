mkdir -p gpurun_out
{
echo "== peer BN test (2 GPUs)"
timeout 300 python -m pytest tests/test_peer_bn_gpu.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error -e "^E " | tail -12
echo "== model tests (1 GPU of the 2)"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s 2>&1 | grep -e "PARITY side\|PARITY graphed" -e passed -e failed -e "^E " | cut -c1-300 | tail -8
} 2>&1 | tee gpurun_out/r2_call17.log
