// capi_multi.cu — one process driving several GPUs: the frame-sharded throughput mode of SURVEY.md §8e as C++ host code.
//
// Frames are independent once calibration is fixed (the reference's per-frame loop, LIB/src/stitcher.cpp:221-313 /
// APP64:724-770, carries nothing from one frame set to the next), so frame f of a sequence goes to device f mod n; every
// device holds its own compositor (tables replicated, ~0.25 GB) and there is no data-path collective.  One host thread per
// device keeps that device's slots full; the threads never touch each other's handles.  bench.py's multi-GPU arm is the
// same partitioning with one PROCESS per GPU (torch.distributed only for the barrier and the max-over-ranks timing).
#include <thread>
#include <vector>

#include "sb_internal.h"

using namespace sb;

struct sb_multi {
    std::vector<sb_compositor *> comps;
    std::vector<int> devices;
    int depth = 1, n_cameras = 0;
};

extern "C" {

int sb_multi_create(const sb_compositor_config *cfg, int n_devices, const int *devices, int depth, sb_multi **out)
{
    SB_ASSERT(cfg && out && n_devices > 0 && devices && depth >= 1 && depth <= 16);
    *out = nullptr;
    sb_multi *m = new sb_multi;
    m->depth = depth; m->n_cameras = cfg->n_cameras;
    for (int k = 0; k < n_devices; ++k) {
        sb_compositor *c = nullptr;
        int rc = sb_compositor_create(cfg, devices[k], &c);
        if (rc == SB_OK) rc = sb_compositor_set_depth(c, depth);
        if (rc != SB_OK) {
            if (c) sb_compositor_destroy(c);
            sb_multi_destroy(m);
            return rc;
        }
        m->comps.push_back(c);
        m->devices.push_back(devices[k]);
    }
    *out = m;
    return SB_OK;
}

int sb_multi_size(const sb_multi *m) { return m ? (int)m->comps.size() : 0; }

sb_compositor *sb_multi_handle(sb_multi *m, int k) { return (m && k >= 0 && k < (int)m->comps.size()) ? m->comps[k] : nullptr; }

int sb_multi_run(sb_multi *m, int n_frames, const sb_image *srcs, sb_image *panos, sb_image *pano_masks)
{
    SB_ASSERT(m && srcs && panos && n_frames >= 0);
    const int n = (int)m->comps.size(), nc = m->n_cameras, depth = m->depth;
    std::vector<int> rcs(n, SB_OK);
    std::vector<std::string> errs(n);
    auto work = [&](int k) {
        sb_compositor *c = m->comps[k];
        std::vector<int> in_flight;                          // slot ids in issue order
        int rc = SB_OK;
        for (int f = k; f < n_frames && rc == SB_OK; f += n) {
            if ((int)in_flight.size() == depth) {
                rc = sb_compositor_wait(c, in_flight.front());
                in_flight.erase(in_flight.begin());
                if (rc != SB_OK) break;
            }
            int slot = -1;
            rc = sb_compositor_enqueue(c, srcs + (size_t)f * nc, &panos[f], pano_masks ? &pano_masks[f] : nullptr, &slot);
            if (rc == SB_OK) in_flight.push_back(slot);
        }
        for (int s : in_flight) {
            const int r2 = sb_compositor_wait(c, s);
            if (rc == SB_OK) rc = r2;
        }
        rcs[k] = rc;
        if (rc != SB_OK) errs[k] = sb_last_error();          // (the error text is per thread)
    };
    std::vector<std::thread> th;
    for (int k = 1; k < n; ++k) th.emplace_back(work, k);
    work(0);
    for (auto &t : th) t.join();
    for (int k = 0; k < n; ++k)
        if (rcs[k] != SB_OK) return fail(rcs[k], "device %d: %s", m->devices[k], errs[k].c_str());
    return SB_OK;
}

void sb_multi_destroy(sb_multi *m)
{
    if (!m) return;
    for (sb_compositor *c : m->comps) sb_compositor_destroy(c);
    delete m;
}

}  // extern "C"
