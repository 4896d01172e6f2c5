// interacting_inst.cu — one (D, U, H) instantiation of the fused InteractingLayer
// kernels; compiled once per shape with -DRS_D= -DRS_U= -DRS_H= so the shapes
// build in parallel (each instantiation is a fully unrolled register kernel).
#include "interacting_kernels.cuh"
#define RS_CAT_(a, b, c, d) a##b##_##c##_##d
#define RS_CAT(a, b, c, d) RS_CAT_(a, b, c, d)
namespace rs {
int RS_CAT(interacting_fwd_, RS_D, RS_U, RS_H)(const IFwdArgs& a) { return fwd_shape<RS_D, RS_U, RS_H>(a); }
int RS_CAT(interacting_bwd_, RS_D, RS_U, RS_H)(const IBwdArgs& a) { return bwd_shape<RS_D, RS_U, RS_H>(a); }
}  // namespace rs
