for d in 0 1; do echo "debug=$d"; B200_SORT_DEBUG=$d python tools/microbench.py sort 2>&1 | grep "sort! 2^30 U32"; done
