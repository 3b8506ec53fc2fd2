// md_kernels.cuh — sm_100a kernels of the moldyn `solve` step loop (f64 throughout).
//
//   K1  k_cell_count / k_scan_* / k_scatter / k_sort_cells / k_reorder   cell binning + counting sort
//   K2  k_build_list                                                     Verlet-skin neighbour list
//   K3  k_force  (+ fused second half-kick and K5 partial sums)          potential.rs:158-216, integrator.rs:47-53
//   K4  k_kick_drift                                                     integrator.rs:28-45, barostat.rs:45-48
//   K5  block→warp deterministic reductions + finalize                   macro_parameters/*.rs, thermostat.rs, barostat.rs
//
// No tensor cores: the pair force is an irregular gather, not a dense contraction.  The streaming kernels are
// HBM-bound (coalesced SoA planes, 128-bit accesses), the force kernel is FP64-pipe / L1-gather bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "md_common.cuh"
#include "md_cells.cuh"
#include "md_lists.cuh"
#include "md_reduce.cuh"
#include "md_force.cuh"
#include "md_integrate.cuh"
#include "md_loop.cuh"
#include "md_tile.cuh"
#include "md_dist_kernels.cuh"
#include "md_multi.cuh"
