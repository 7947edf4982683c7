/* Empty stand-in for <CL/cl_ext.h>, see CL/cl.h in this directory. */
#ifndef PSM_SHIM_CL_EXT_H
#define PSM_SHIM_CL_EXT_H
#endif
