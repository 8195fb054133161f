for s in 4 6 8; do for a in 2 3 4; do echo "search $s accum $a"; OPB_ICP_SEARCH_CTAS=$s OPB_ICP_ACCUM_CTAS=$a python scripts/bench_icp.py 3 | tail -1; done; done
