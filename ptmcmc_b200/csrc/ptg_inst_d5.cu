// generated: thread-per-chain kernels for dim=5
#include "ptg_inst.cuh"
PTG_INSTANTIATE(5)
