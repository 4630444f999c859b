set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -8
timeout 120 python tools/fused_bench.py 8192 10
AVF_LIB_OVERRIDE=build/libavf_prof.so timeout 120 python tools/fused_phases.py 8192
