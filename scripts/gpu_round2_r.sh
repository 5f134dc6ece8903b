# peer-memory exchange: one-GPU tests (processes sharing cuda:0)
mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_cl_sharded.py -x -q -s 2>&1 | tail -15 | tee gpurun_out/r2r/peer_tests.txt
