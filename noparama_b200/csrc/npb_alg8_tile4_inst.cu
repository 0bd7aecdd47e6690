// npb_alg8_tile4_inst.cu -- one explicit instantiation of the four-chains-per-CTA sweep kernel launcher; compiled with
// -DNPB_INST_D=<D> (see Makefile)
#include "npb_alg8_tile4.cuh"
template npb_status npb_launch_alg8_tile4<NPB_INST_D>(npb_chains *, const SweepArgs &);
template npb_status npb_launch_aux_keys<NPB_INST_D>(npb_chains *, const SweepArgs &);
template npb_status npb_launch_tile4_probe<NPB_INST_D>(npb_chains *, const SweepArgs &, int, const int32_t *, float *);
