// pipeline.cu -- one host call per streamed frame.
//
// radnerf_b200.stream.FrameStreamer keeps several frames in flight (lanes); in steady state everything a frame needs is a fixed
// set of buffers, streams, events and one captured graph per lane.  Issuing that from Python costs ~0.2 ms per frame (torch
// stream contexts, current-stream queries, ctypes marshalling of four separate calls), which is MORE than the device needs
// for a ray-sharded frame on 4-8 GPUs.  rn_lane_submit_frame issues the whole choreography from C:
//
//   lane stream : copy the input block in -> generate rays -> record ev_in
//   cond stream : wait ev_in, wait ev_done (lane's previous frame) -> conditioning kernel -> record ev_cond
//   lane stream : wait ev_cond -> launch the lane's frame graph -> [scatter image rows to the peers]          (phase 1)
//   lane stream : [wait ev_delivered] -> stage the image -> record ev_staged ; copy stream: wait, D2H, record ev_delivered
//                 -> record ev_done                                                                           (phase 2)
//
// The split in two phases exists because the multi-GPU image assembly needs a cross-rank barrier between the scatter and
// the staging copy, and that barrier is torch's symmetric-memory barrier, issued by the caller.
#include <cstring>
#include "common.cuh"
#include "../../include/radnerf_b200.h"

using namespace rn;

#define RN_CU(call)                                                                      \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            rn::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return (int)e_;                                                              \
        }                                                                                \
    } while (0)

extern "C" int rn_event_create(void** ev) {
    RN_REQUIRE(ev, "null pointer");
    cudaEvent_t e;
    RN_CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *ev = (void*)e;
    return RN_OK;
}

extern "C" int rn_event_destroy(void* ev) {
    if (ev) RN_CU(cudaEventDestroy((cudaEvent_t)ev));
    return RN_OK;
}

extern "C" int rn_event_synchronize(void* ev) {
    RN_REQUIRE(ev, "null event");
    RN_CU(cudaEventSynchronize((cudaEvent_t)ev));
    return RN_OK;
}

extern "C" int rn_stream_wait_event(void* stream, void* ev) {
    RN_REQUIRE(ev, "null event");
    RN_CU(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
    return RN_OK;
}

extern "C" int rn_lane_submit_frame(const rn_lane_submit* s) {
    RN_REQUIRE(s, "null descriptor");
    cudaStream_t ls = (cudaStream_t)s->lane_stream, cs = (cudaStream_t)s->cond_stream, xs = (cudaStream_t)s->copy_stream;
    if (s->phase & 1u) {
        RN_REQUIRE(s->ev_in && s->ev_cond && s->ev_done && s->graph_exec && s->cond, "phase 1 needs events, graph and conditioning");
        if (s->packed_bytes) RN_CU(cudaMemcpyAsync(s->flat_dst, s->packed_src, s->packed_bytes, cudaMemcpyDefault, ls));
        if (s->n_rays) {
            int rc = rn_get_rays(s->pose, s->fx, s->fy, s->cx, s->cy, s->H, s->W, s->pixel_ids, s->n_rays, s->rays_o, s->rays_d, ls);
            if (rc) return rc;
        }
        // conditioning in two launches: the audio nets of THIS frame on the lane's own stream (concurrent across lanes; the lane's
        // previous frame, which read the hoisted-term vectors, precedes it in stream order), then only the lip-smoothing EMA + the
        // hoisted products in frame order on the conditioning stream
        rn_conditioning_desc cd = *s->cond;
        const bool split = cd.auds != nullptr && cd.reserved == 0;
        if (split) {
            cd.reserved = 1;
            int rc1 = rn_frame_conditioning(&cd, ls);
            if (rc1) return rc1;
            cd.reserved = 2;
        }
        RN_CU(cudaEventRecord((cudaEvent_t)s->ev_in, ls));
        RN_CU(cudaStreamWaitEvent(cs, (cudaEvent_t)s->ev_in, 0));     // the conditioning reads the input block / the parked raw code
        RN_CU(cudaStreamWaitEvent(cs, (cudaEvent_t)s->ev_done, 0));   // lane's previous frame no longer reads its hoisted-term vectors
        int rc = rn_frame_conditioning(&cd, cs);
        if (rc) return rc;
        RN_CU(cudaEventRecord((cudaEvent_t)s->ev_cond, cs));
        RN_CU(cudaStreamWaitEvent(ls, (cudaEvent_t)s->ev_cond, 0));
        RN_CU(cudaGraphLaunch((cudaGraphExec_t)s->graph_exec, ls));
        rn_note_graph_replay(s->graph_kernels);
        if (s->peers && s->world > 1) {
            rc = s->ctrl_peers ? rn_scatter_rows_to_root(s->image_local, s->ids, s->n_local, s->run_pixels, s->peers, s->ctrl_peers, s->world, s->rank,
                                                         s->root, s->slot, s->frame_seq, ls)
                               : rn_scatter_rows_to_peers(s->image_local, s->ids, s->n_local, s->run_pixels, s->peers, s->world, ls);
            if (rc) return rc;
        }
    }
    if (s->phase & 2u) {
        RN_REQUIRE(s->ev_done, "phase 2 needs ev_done");
        const bool exchange = s->ctrl_peers && s->world > 1;
        if (exchange && s->rank == s->root) {
            // the root assembles: wait for the other ranks' rows (arrival counter), stage, release the buffer (consumed flags) -- one kernel
            const bool deliver = s->host_dst != nullptr;
            if (deliver) {
                RN_REQUIRE(s->ev_staged && s->ev_delivered && s->stage_src && s->stage_dst, "delivery needs staging buffers and events");
                RN_CU(cudaStreamWaitEvent(ls, (cudaEvent_t)s->ev_delivered, 0));   // the copy that last read this staging slot has drained
            }
            int rc = rn_stage_frame_at_root((const float*)s->stage_src, deliver ? s->stage_dst : nullptr, s->image_bytes / 4, s->to_uint8, s->ctrl_peers,
                                            s->world, s->root, s->slot, s->frame_seq, s->scatter_ctas, s->ticket, ls);
            if (rc) return rc;
            if (deliver) {
                RN_CU(cudaEventRecord((cudaEvent_t)s->ev_staged, ls));
                RN_CU(cudaStreamWaitEvent(xs, (cudaEvent_t)s->ev_staged, 0));
                RN_CU(cudaMemcpyAsync(s->host_dst, s->stage_dst, s->to_uint8 ? s->image_bytes / 4 : s->image_bytes, cudaMemcpyDeviceToHost, xs));
                RN_CU(cudaEventRecord((cudaEvent_t)s->ev_delivered, xs));
            }
        } else if (s->host_dst && !exchange) {
            RN_REQUIRE(s->ev_staged && s->ev_delivered && s->stage_src && s->stage_dst, "delivery needs staging buffers and events");
            RN_CU(cudaStreamWaitEvent(ls, (cudaEvent_t)s->ev_delivered, 0));   // the copy that last read this staging slot has drained
            uint64_t out_bytes = s->image_bytes;
            if (s->to_uint8) {   // output stage on the device: a quarter of the bytes cross PCIe
                out_bytes = s->image_bytes / 4;
                int rc = rn_image_to_uint8((const float*)s->stage_src, (uint8_t*)s->stage_dst, out_bytes, ls);
                if (rc) return rc;
            } else {
                RN_CU(cudaMemcpyAsync(s->stage_dst, s->stage_src, s->image_bytes, cudaMemcpyDeviceToDevice, ls));
            }
            RN_CU(cudaEventRecord((cudaEvent_t)s->ev_staged, ls));
            RN_CU(cudaStreamWaitEvent(xs, (cudaEvent_t)s->ev_staged, 0));
            RN_CU(cudaMemcpyAsync(s->host_dst, s->stage_dst, out_bytes, cudaMemcpyDeviceToHost, xs));
            RN_CU(cudaEventRecord((cudaEvent_t)s->ev_delivered, xs));
        }
        RN_CU(cudaEventRecord((cudaEvent_t)s->ev_done, ls));
    }
    return RN_OK;
}

// sizes of the public descriptor structs, so that a binding (radnerf_b200/frame.py, stream.py: ctypes.Structure mirrors) can
// verify its layout against the library it loaded
extern "C" uint32_t rn_sizeof(const char* name) {
    if (!name) return 0;
    const struct { const char* n; uint32_t s; } table[] = {
        {"rn_grid_table", (uint32_t)sizeof(rn_grid_table)},           {"rn_conditioning_desc", (uint32_t)sizeof(rn_conditioning_desc)},
        {"rn_frame_head_desc", (uint32_t)sizeof(rn_frame_head_desc)}, {"rn_frame_torso_desc", (uint32_t)sizeof(rn_frame_torso_desc)},
        {"rn_lane_submit", (uint32_t)sizeof(rn_lane_submit)},
        {"rn_adam_tensor", (uint32_t)sizeof(rn_adam_tensor)},         {"rn_adam_group", (uint32_t)sizeof(rn_adam_group)},
        {"rn_ring_windows", (uint32_t)sizeof(rn_ring_windows)},
    };
    for (const auto& e : table)
        if (strcmp(e.n, name) == 0) return e.s;
    return 0;
}
