/* Minimal stand-in for the Khronos <CL/cl.h>, written for this repository.
 *
 * TEST INFRASTRUCTURE ONLY. It exists so that the reference's CPU_RUN build
 * (which never calls OpenCL but still includes clinterface.h / random.h /
 * data_analysis.h, all of which name cl_* types) can be compiled in a
 * container that has no OpenCL headers. Only scalar typedefs, opaque handle
 * typedefs and the vector unions the reference headers mention are declared;
 * no entry points.
 */
#ifndef QCDGPU_B200_STUB_CL_H
#define QCDGPU_B200_STUB_CL_H

#include <stddef.h>
#include <stdint.h>

typedef int8_t   cl_char;
typedef uint8_t  cl_uchar;
typedef int16_t  cl_short;
typedef uint16_t cl_ushort;
typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef int64_t  cl_long;
typedef uint64_t cl_ulong;
typedef float    cl_float;
typedef double   cl_double;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_uint  cl_device_info;
typedef cl_uint  cl_platform_info;

typedef struct _stub_cl_platform_id*   cl_platform_id;
typedef struct _stub_cl_device_id*     cl_device_id;
typedef struct _stub_cl_context*       cl_context;
typedef struct _stub_cl_command_queue* cl_command_queue;
typedef struct _stub_cl_mem*           cl_mem;
typedef struct _stub_cl_program*       cl_program;
typedef struct _stub_cl_kernel*        cl_kernel;
typedef struct _stub_cl_event*         cl_event;

typedef union { cl_float  s[2]; struct { cl_float  x, y; }; }       cl_float2;
typedef union { cl_float  s[4]; struct { cl_float  x, y, z, w; }; } cl_float4;
typedef union { cl_double s[2]; struct { cl_double x, y; }; }       cl_double2;
typedef union { cl_double s[4]; struct { cl_double x, y, z, w; }; } cl_double4;
typedef union { cl_uint   s[2]; struct { cl_uint   x, y; }; }       cl_uint2;
typedef union { cl_uint   s[4]; struct { cl_uint   x, y, z, w; }; } cl_uint4;
typedef union { cl_int    s[4]; struct { cl_int    x, y, z, w; }; } cl_int4;

#define CL_SUCCESS 0
#define CL_TRUE    1
#define CL_FALSE   0

#endif
