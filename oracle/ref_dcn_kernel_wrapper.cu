// Compiles the reference's OWN deformable-conv CUDA kernels, where they lie under /root/reference, for sm_100a.
// Nothing of the reference is copied: this file only pre-includes the headers the reference source asks for and then
// #includes it by path (REF_DCN_KERNEL_CU is passed by oracle/build_ref.py).  The single macro below is what lets the
// 2019-era source build against torch 2.x: its six AT_DISPATCH_FLOATING_TYPES_AND_HALF(x.type(), ...) sites
// (dcn_deform_conv_cuda_kernel.cu:258, 352, 450, 780, 812, 845) need x.scalar_type() today.
// TEST INFRASTRUCTURE: the product never loads the resulting module.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <THC/THCAtomics.cuh>
#include <stdio.h>
#include <math.h>
#include <float.h>
#define type scalar_type
#include REF_DCN_KERNEL_CU
