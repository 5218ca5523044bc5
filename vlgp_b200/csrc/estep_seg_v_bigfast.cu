// Segment E-step variant: factors of up to 32 columns (omega near its upper bound), all channels Poisson with uint8 counts
#include "estep_seg_impl.cuh"
VLGP_DEFINE_SEG_VARIANT(4, 1)
