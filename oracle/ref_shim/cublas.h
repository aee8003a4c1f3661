/* Include-path shim used ONLY by oracle/build_ref.sh when compiling the unmodified
 * reference extension from /root/reference/extension (test infrastructure, not product).
 * The reference's common.h:6 includes the legacy <cublas.h>, which clashes with the
 * <cublas_v2.h> that ATen/cuda/CUDAContext.h pulls in ("#error It is an error to include
 * both...").  Putting this directory first on the include path resolves <cublas.h> to the
 * v2 header instead, so no reference source has to be copied or edited. */
#pragma once
#include <cublas_v2.h>
