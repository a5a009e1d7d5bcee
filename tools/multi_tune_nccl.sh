#!/bin/bash
# NCCL point-to-point knobs for the root-egress-bound scatter of qppvm_multi_solve_batch (8 GPUs, one process)
S="1.0,16384 1.0,32768"
TUNE_TAG=default python tools/multi_tune.py 8 "$S" 2>&1 | grep "G=8"
TUNE_TAG=memcpy NCCL_P2P_USE_CUDA_MEMCPY=1 python tools/multi_tune.py 8 "$S" 2>&1 | grep "G=8"
TUNE_TAG=minch4 NCCL_MIN_P2P_NCHANNELS=4 python tools/multi_tune.py 8 "$S" 2>&1 | grep "G=8"
TUNE_TAG=minch8 NCCL_MIN_P2P_NCHANNELS=8 NCCL_MAX_P2P_NCHANNELS=64 python tools/multi_tune.py 8 "$S" 2>&1 | grep "G=8"
TUNE_TAG=chunk2M NCCL_P2P_NVL_CHUNKSIZE=2097152 python tools/multi_tune.py 8 "$S" 2>&1 | grep "G=8"
