// Row class 40 of the register-resident fused solve (kernel template: pop_fused2.cuh).
#include "pop_fused2.cuh"
F2_DEFINE_CLASS(40)
