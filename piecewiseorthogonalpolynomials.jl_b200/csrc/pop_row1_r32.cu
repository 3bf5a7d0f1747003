// Chain class 32 of the one-pass row half step (kernel template: pop_row1.cuh).
#include "pop_row1.cuh"
R1_DEFINE_CLASS(32)
