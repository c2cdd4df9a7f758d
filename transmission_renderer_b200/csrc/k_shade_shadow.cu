// k_shade_shadow.cu — the shading kernels instantiated with ray-queried shadows, and the shadow pass that feeds them.
// Everything lives in k_shade.cu; this translation unit only selects the TR_SHADE_SHADOW=1 half of it so that the two
// halves compile side by side and the instantiations without ray queries stay exactly as they were.
#define TR_SHADE_SHADOW 1
#include "k_shade.cu"
