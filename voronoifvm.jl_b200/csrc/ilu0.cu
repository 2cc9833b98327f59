// K10: ILU(0) preconditioner -- placeholder until the level-scheduled kernels land.
#include "vfvm_internal.h"

void vfvm_ilu0_setup(vfvm_handle* h) { throw std::string("ILU0 preconditioner is not built yet; use Jacobi or block-Jacobi"); }
void vfvm_ilu0_apply(vfvm_handle* h, const double* in, double* out) { throw std::string("ILU0 preconditioner is not built yet"); }
