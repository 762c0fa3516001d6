#!/bin/bash
# k_pool tuning sweep (one gpurun call): RNG inline vs out-of-line, pool geometry, phase thresholds
export AB_ONLY=${AB_ONLY:-book1_final,mesh871k}
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_pool.py time 2>&1 | grep pool; }
echo "== auto"; timeout 120 python tools/ab_pool.py time 2>&1 | grep auto
run X=1
run RTB200_LIB=$PWD/ray_tracing_series_rust_b200/librtb200_inl.so
run RTB200_POOL_SLOTS=96 RTB200_POOL_OCC=5
run RTB200_POOL_SLOTS=96
run RTB200_POOL_REFILL=4
run RTB200_POOL_REFILL=12
run RTB200_POOL_REFILL=16
run RTB200_POOL_LOW=16
run RTB200_POOL_LOW=20
run RTB200_POOL_LOW=28
run RTB200_POOL_LOW=28 RTB200_POOL_REFILL=4
