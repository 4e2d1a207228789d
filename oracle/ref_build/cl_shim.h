/*
 * cl_shim.h -- the handful of OpenCL C built-ins /root/reference/opencv/createMap.cl uses,
 * spelled for gcc, so that the reference's kernel source compiles UNMODIFIED as plain C
 * (createmap_driver.c: #include "cl_shim.h", then #include of the .cl file).  TEST INFRASTRUCTURE ONLY.
 *
 * What is fixed here because OpenCL C leaves it to the implementation:
 *   - dot(float3, float3): ((a0*b0 + a1*b1) + a2*b2), every operation rounded once to fp32
 *     (compile with -ffp-contract=off, no -march flags), the order oracle/create_map_ref.c documents;
 *   - length(float2): sqrtf(x*x + y*y);
 *   - atan(float): the float overload (atanf from glibc; OpenCL allows 5 ulp, glibc is < 1 ulp);
 *   - get_global_id(): read from thread-local variables the driver (createmap_driver.c) sets per
 *     work-item; mad24(): integer multiply-add.
 * Vector types are gcc vector extensions: brace initialisers and [] subscripts work as the
 * kernel writes them (createMap.cl:15-35); float3 occupies four lanes like OpenCL's.
 */
#ifndef VAW_CL_SHIM_H
#define VAW_CL_SHIM_H
#include <math.h>
#include <stddef.h>

#define __kernel static inline
#define __global

typedef float float2 __attribute__((vector_size(8)));
typedef float float3 __attribute__((vector_size(16)));

extern __thread size_t vaw_cl_global_id[3];
static inline size_t get_global_id(unsigned dim) { return vaw_cl_global_id[dim]; }

static inline float vaw_cl_dot3(float3 a, float3 b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline float vaw_cl_length2(float2 v) { return sqrtf(v[0] * v[0] + v[1] * v[1]); }
static inline int vaw_cl_mad24(int a, int b, int c) { return a * b + c; }

#define dot(a, b) vaw_cl_dot3((a), (b))
#define length(v) vaw_cl_length2((v))
#define mad24(a, b, c) vaw_cl_mad24((int)(a), (int)(b), (int)(c))
#define atan(x) atanf((x))

#endif
