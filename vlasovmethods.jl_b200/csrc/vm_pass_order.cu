// vm_pass_order.cu -- instantiations of the fused particle pass for ONE spline order (VM_PASS_ORDER).
#include "vm_pass.cuh"
#include "vm_pass_bq.cuh"

#ifndef VM_PASS_ORDER
#error "compile with -DVM_PASS_ORDER=<2..6>"
#endif
#define VM_CAT2(a, b) a##b
#define VM_CAT(a, b) VM_CAT2(a, b)

void VM_CAT(vm_launch_pass_k, VM_PASS_ORDER)(vm_ctx* ctx, int mode, const DepositPlan& pl, double* x, double* v, const double* w,
                                             const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    switch (mode) {
        case MODE_DEPOSIT: launch_pass_var<VM_PASS_ORDER, MODE_DEPOSIT>(ctx, pl, x, v, w, dcoef, out, P, F); break;
        case MODE_PUSH_DEPOSIT: launch_pass_var<VM_PASS_ORDER, MODE_PUSH_DEPOSIT>(ctx, pl, x, v, w, dcoef, out, P, F); break;
        default: launch_pass_var<VM_PASS_ORDER, MODE_DRIFT_DEPOSIT>(ctx, pl, x, v, w, dcoef, out, P, F); break;
    }
}

// bank-sorted deposition for large meshes (vm_pass_bq.cuh)
void VM_CAT(vm_launch_pass_bq_k, VM_PASS_ORDER)(vm_ctx* ctx, int mode, const BqPlan& bp, double* x, double* v, const double* w,
                                                const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    switch (mode) {
        case MODE_DEPOSIT: launch_bq<VM_PASS_ORDER, MODE_DEPOSIT>(ctx, bp, x, v, w, dcoef, out, P, F); break;
        case MODE_PUSH_DEPOSIT: launch_bq<VM_PASS_ORDER, MODE_PUSH_DEPOSIT>(ctx, bp, x, v, w, dcoef, out, P, F); break;
        default: launch_bq<VM_PASS_ORDER, MODE_DRIFT_DEPOSIT>(ctx, bp, x, v, w, dcoef, out, P, F); break;
    }
}
