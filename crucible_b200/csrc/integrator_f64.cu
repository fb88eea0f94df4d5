// integrator_f64.cu — the bit-faithful instantiation (R = double).
// This translation unit MUST be compiled with -fmad=false: the reference (rustc) never contracts
// a*b+c into an FMA, so neither may we (hard part 3).
#include <cstring>

#include "integrator.cuh"

namespace crb {
template int render_impl<double>(const SceneDeviceData&, Workspace&, const CrCamera&, const CrRenderOpts&, void*, void*, int,
                                 cudaStream_t, CrStats*, std::string&);
template int trace_batch_impl<double>(const SceneDeviceData&, const double*, size_t, double, double, CrHit*, uint32_t*, uint32_t*, int,
                                      uint32_t*, cudaStream_t, std::string&);
}  // namespace crb
