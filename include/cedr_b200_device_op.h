/* Device-callable DeviceOp methods over the POD view exported by
 * cedr_b200_get_device_op(): the reference's
 *   QLT<ES>::DeviceOp::set_rhom / set_Qm / get_Qm   (cedr_qlt_inl.hpp:14-66)
 *   CAAS<ES>::DeviceOp::set_rhom / set_Qm / get_Qm  (cedr_caas_inl.hpp:13-42)
 * Include from CUDA translation units; copy the struct by value into kernels.
 * Concurrent calls for distinct (lclcellidx, tracer_idx) are safe; set_rhom must
 * precede set_Qm for consistent-only tracers (cedr_cdr.hpp:80).
 */
#ifndef CEDR_B200_DEVICE_OP_H
#define CEDR_B200_DEVICE_OP_H

#include "cedr_b200.h"

#ifdef __CUDACC__

__device__ __forceinline__ void
cedr_b200_op_set_rhom (const cedr_b200_device_op& op, int lclcellidx, int /*rhomidx*/,
                       double rhom) {
  op.in[lclcellidx] = rhom;
}

__device__ __forceinline__ void
cedr_b200_op_set_Qm (const cedr_b200_device_op& op, int lclcellidx, int tracer_idx,
                     double Qm, double Qm_min, double Qm_max,
                     double Qm_prev = __longlong_as_double(0x7ff0000000000000LL)) {
  const int pt = op.trcr_prob[tracer_idx];
  double* bd = op.in + (int64_t) op.trcr_row[tracer_idx]*op.ld + lclcellidx;
  int next;
  if (pt & CEDR_B200_SHAPEPRESERVE) {
    bd[0] = Qm_min;
    bd[op.ld] = Qm;
    bd[2*op.ld] = Qm_max;
    next = 3;
  } else if (pt & CEDR_B200_CONSISTENT) {
    const double rhom = op.in[lclcellidx];
    bd[0] = Qm_min/rhom;
    bd[op.ld] = Qm;
    bd[2*op.ld] = Qm_max/rhom;
    next = 3;
  } else {
    bd[0] = Qm;
    next = 1;
  }
  if ((pt & CEDR_B200_CONSERVE) || (op.is_caas && op.reserved)) {
    /* The reference aborts the kernel when a conserving tracer is set without
     * Qm_prev (cedr_qlt_inl.hpp:53-55). */
    if ((pt & CEDR_B200_CONSERVE) && Qm_prev == __longlong_as_double(0x7ff0000000000000LL))
      __trap();
    bd[next*op.ld] = Qm_prev;
  }
}

__device__ __forceinline__ double
cedr_b200_op_get_Qm (const cedr_b200_device_op& op, int lclcellidx, int tracer_idx) {
  if (op.is_caas)
    return op.in[((int64_t) op.trcr_row[tracer_idx] + 1)*op.ld + lclcellidx];
  return op.out[(int64_t) tracer_idx*op.ld + lclcellidx];
}

#endif /* __CUDACC__ */
#endif
