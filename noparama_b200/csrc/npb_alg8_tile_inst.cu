// npb_alg8_tile_inst.cu -- one explicit instantiation of the D >= 4 sweep kernel launcher; compiled with
// -DNPB_INST_D=<D> -DNPB_INST_KMAX=<32|64> (see Makefile)
#include "npb_alg8_tile.cuh"
template npb_status npb_launch_alg8_tile<NPB_INST_D, NPB_INST_KMAX>(npb_chains *, const SweepArgs &);
