// cells.cu — sorted cell-list path (placeholder until the first all-pairs GPU run is green).
#include "ljmd_internal.cuh"
namespace ljmd {
struct Cells {};
int  cells_create(ljmd_handle*) { set_error("cell-list path not built yet"); return LJMD_E_UNSUPPORTED; }
void cells_destroy(ljmd_handle*) {}
int  cells_run(ljmd_handle*, const float2*, const float2*, float2*, float2*, float2*, float*, const RunCtl&) { return LJMD_E_UNSUPPORTED; }
int  cells_geometry(ljmd_handle*, int*, float*, float*) { return LJMD_E_UNSUPPORTED; }
int  cells_assign(ljmd_handle*, const float2*, int*, int*) { return LJMD_E_UNSUPPORTED; }
int  cells_neighbor_count(ljmd_handle*, const float2*, float, int*) { return LJMD_E_UNSUPPORTED; }
long long cells_last_rebuilds(ljmd_handle*) { return 0; }
}
