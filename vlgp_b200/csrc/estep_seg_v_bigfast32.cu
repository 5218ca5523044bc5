// Segment E-step variant: single-precision rate passes (FAST = 2), factors of up to 32 columns
#include "estep_seg_impl.cuh"
VLGP_DEFINE_SEG_VARIANT(4, 2)
