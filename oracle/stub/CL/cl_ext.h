/* Stand-in for <CL/cl_ext.h>: nothing from it is needed by the CPU_RUN build. */
#ifndef QCDGPU_B200_STUB_CL_EXT_H
#define QCDGPU_B200_STUB_CL_EXT_H
#include "cl.h"
#endif
