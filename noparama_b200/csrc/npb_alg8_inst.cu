// npb_alg8_inst.cu -- one explicit instantiation of the sweep kernel launcher; compiled with
// -DNPB_INST_D=<D> -DNPB_INST_SPL=<register levels> (see Makefile)
#include "npb_alg8_kernel.cuh"
template npb_status npb_launch_alg8_reg<NPB_INST_D, NPB_INST_SPL>(npb_chains *, const SweepArgs &);
