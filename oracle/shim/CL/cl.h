/* Empty stand-in for <CL/cl.h> (TEST INFRASTRUCTURE, oracle/_ref build only).
 * /root/reference/include/ComFunc.h:33 includes it unconditionally; the three reference sources
 * compiled into oracle/_ref (src/CVC.cpp, src/CVF.cpp, src/DispSel.cpp) use nothing from OpenCL. */
#ifndef PSM_SHIM_CL_H
#define PSM_SHIM_CL_H
#endif
