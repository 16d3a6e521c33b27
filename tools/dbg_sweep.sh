for d in 0 1 2 3 4; do echo "== NS2_GEMM_DEBUG=$d"; NS2_GEMM_DEBUG=$d python tools/prof_kernels.py sweep 2>&1 | grep -E "K=512|N=1024" ; done
