// Instantiation of the segment E-step kernel: NBMAX = 2, FAST = 1 (see estep_seg_impl.cuh).
#include "estep_seg_impl.cuh"
VLGP_DEFINE_SEG_VARIANT(2, 1)
