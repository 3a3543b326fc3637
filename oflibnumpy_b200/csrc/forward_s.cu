// Source-referenced (forward) resampling -- placeholder entry points until the rasteriser lands.
#include "ofk_common.cuh"
using namespace ofk;

extern "C" size_t ofk_forward_s_workspace(int N, int H, int W) { return (size_t)N * H * W * 8; }

extern "C" int ofk_forward_s(const float*, int, const float*, float, const uint8_t*, const uint8_t*, float*, uint8_t*,
                             int, int, int, void*, size_t, ofk_stream_t) {
    set_error("ofk_forward_s: not implemented in this build");
    return OFK_EUNSUPPORTED;
}
