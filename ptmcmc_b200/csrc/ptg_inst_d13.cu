// generated: thread-per-chain kernels for dim=13
#include "ptg_inst.cuh"
PTG_INSTANTIATE(13)
