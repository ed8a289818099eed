// od_history.cuh -- the run loop's output buffer on the device.
//
// OpenDriftSimulation.state_to_buffer (opendrift/models/basemodel/__init__.py:2384-2499) writes, at every output step,
// lon / lat / z / status of the active elements into the `[trajectory, time]` arrays of the result, addressed by element
// ID (`self.result[var][ID_ind, step]`, float32 for the positions).  Here the same arrays live in HBM for
// `export_buffer_length` output steps; one thread per active element scatters its four values into column `col`, and
// the host reads the block back once per `export_buffer_length` steps instead of four arrays per step.
#pragma once
#include <stdint.h>

namespace od {

struct HistoryParams {
    int64_t n;               // active elements
    int64_t n_total;         // trajectories (rows of the buffers)
    int32_t col, ncols;      // output column to write, columns per row
    int32_t z_f64, pad_;
    const int32_t* ids;      // element ID (row index), [n]
    const double* lon;
    const double* lat;
    const void* z;           // float32, or float64 when z_f64
    const int32_t* status;
    float* blon;             // [n_total][ncols]
    float* blat;
    float* bz;
    int32_t* bstatus;
};

#if defined(__CUDACC__)
#define OD_HIST_HD __host__ __device__ __forceinline__
#else
#define OD_HIST_HD static inline
#endif

OD_HIST_HD void history_scatter_one(const HistoryParams& p, int64_t i) {
    const int64_t id = p.ids[i];
    if (id < 0 || id >= p.n_total) return;
    const int64_t o = id * p.ncols + p.col;
    p.blon[o] = (float)p.lon[i];
    p.blat[o] = (float)p.lat[i];
    p.bz[o] = p.z_f64 ? (float)((const double*)p.z)[i] : ((const float*)p.z)[i];
    p.bstatus[o] = p.status[i];
}

}  // namespace od
