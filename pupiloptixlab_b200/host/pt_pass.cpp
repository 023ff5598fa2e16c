#include "pt_pass.h"

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <vector>

namespace Pupil::pt {
namespace {
PTPass *g_active_pass = nullptr; // events outlive passes: only the live pass reacts
}

PTPass::PTPass(std::string_view name) noexcept : Pass(name) {
    g_active_pass = this;
    BindingEventCallback();
}

PTPass::~PTPass() noexcept {
    if (g_active_pass == this) g_active_pass = nullptr;
}

void PTPass::OnRun() noexcept {
    if (!m_world || !m_params.accum_buffer) return;
    if (m_dirty) { // pt_pass.cpp:40-49
        m_params.config.max_depth = m_max_depth;
        m_params.config.accumulated_flag = m_accumulated_flag;
        m_params.sample_cnt = 0;
        m_params.random_seed = m_first_seed;
        m_shard_step = 0, m_shard_total_spp = 0; // sums start over: accumulate = 2 with sample_cnt = 0 does not read the buffer
        m_dirty = false;
    }
    m_params.handle = m_world->GetSceneHandle(); // refreshes BVH / camera / emitters when they changed
    if (!m_params.handle) return;

    if (m_comm) {
        OnRunSharded();
        return;
    }
    pb2_launch_params lp{};
    lp.max_depth = m_params.config.max_depth;
    lp.accumulate = m_sum_mode ? 2u : (m_params.config.accumulated_flag ? 1u : 0u);
    lp.width = m_params.config.frame.width, lp.height = m_params.config.frame.height;
    lp.random_seed = m_params.random_seed, lp.seed_stride = m_seed_stride;
    lp.sample_cnt = m_params.sample_cnt, lp.n_frames = m_frames_per_run;
    lp.accum_buffer = m_params.accum_buffer, lp.frame_buffer = m_params.frame_buffer;
    lp.normal_buffer = m_params.normal_buffer, lp.albedo_buffer = m_params.albedo_buffer, lp.test_buffer = m_params.test;
    if (pb2_render(m_params.handle, &lp) != PB2_OK || pb2_synchronize(m_params.handle) != PB2_OK) {
        Log::Error("pb2_render: %s", pb2_last_error());
        return;
    }
    m_params.sample_cnt += (m_params.config.accumulated_flag ? 1u : 0u) * m_frames_per_run; // :55
    m_params.random_seed += m_frames_per_run * m_seed_stride;                                // :56
}

// One progressive step of a sharded render: this rank's seeds into the sum buffer, then the asynchronous reduction.
void PTPass::OnRunSharded() noexcept {
    uint32_t first = 0, stride = 1, mine = 0, total = 0;
    if (pb2_shard_plan(m_rank, m_n_ranks, m_shard_step, m_frames_per_run, m_strong ? 1 : 0, &first, &stride, &mine, &total) != PB2_OK) {
        Log::Error("pb2_shard_plan: %s", pb2_last_error());
        return;
    }
    pb2_launch_params lp{};
    lp.max_depth = m_params.config.max_depth;
    lp.accumulate = 2u;
    lp.width = m_params.config.frame.width, lp.height = m_params.config.frame.height;
    lp.random_seed = m_first_seed + first, lp.seed_stride = stride;
    lp.sample_cnt = m_params.sample_cnt, lp.n_frames = mine;
    lp.accum_buffer = m_params.accum_buffer, lp.frame_buffer = nullptr;
    lp.normal_buffer = m_params.normal_buffer, lp.albedo_buffer = m_params.albedo_buffer, lp.test_buffer = m_params.test;
    if (mine > 0) {
        if (pb2_render(m_params.handle, &lp) != PB2_OK) {
            Log::Error("pb2_render: %s", pb2_last_error());
            return;
        }
    } else if (m_params.sample_cnt == 0) { // more ranks than samples: this rank contributes zeros
        Synchronize();
        pb2_memset(m_params.accum_buffer, 0, m_output_pixel_num * sizeof(float) * 4);
    }
    m_params.sample_cnt += mine, m_shard_total_spp += total, ++m_shard_step;
    m_params.random_seed = m_first_seed + m_shard_step * total;
    if (pb2_comm_reduce_frames(m_comm, m_params.handle, m_params.accum_buffer, m_params.frame_buffer, m_output_pixel_num, m_shard_total_spp, m_reduce_mode, 0) != PB2_OK)
        Log::Error("pb2_comm_reduce_frames: %s", pb2_last_error());
}
void PTPass::SetShard(pb2_comm *comm, int rank, int world, bool strong, int reduce_mode) noexcept {
    Synchronize();
    m_comm = comm, m_rank = rank, m_n_ranks = world > 0 ? world : 1, m_strong = strong, m_reduce_mode = reduce_mode;
    m_dirty = true;
}
void PTPass::Synchronize() noexcept {
    if (m_params.handle) pb2_synchronize(m_params.handle);
    if (m_comm) pb2_comm_synchronize(m_comm);
}

void PTPass::SetScene(world::World *world) noexcept {
    if (world == nullptr) { // a scene load failed (System::AfterSceneLoadFailed): nothing to render until the next SetScene
        m_world = nullptr;
        m_params = LaunchParams{};
        m_output_pixel_num = 0;
        m_dirty = true;
        return;
    }
    m_world = world;
    m_params.config.frame.width = world->scene->sensor.film.w;
    m_params.config.frame.height = world->scene->sensor.film.h;
    m_params.config.max_depth = world->scene->integrator.max_depth;
    m_params.config.accumulated_flag = true;
    m_max_depth = m_params.config.max_depth, m_accumulated_flag = true;
    m_params.random_seed = 0, m_params.sample_cnt = 0;
    m_output_pixel_num = static_cast<size_t>(m_params.config.frame.width) * m_params.config.frame.height;

    auto *buf_mngr = util::Singleton<BufferManager>::instance();
    Buffer *final_result = buf_mngr->GetBuffer(buf_mngr->DEFAULT_FINAL_RESULT_BUFFER_NAME);
    m_params.frame_buffer = final_result ? final_result->cuda_ptr : nullptr;
    BufferDesc desc{};
    desc.width = m_params.config.frame.width, desc.height = m_params.config.frame.height;
    auto alloc = [&](const char *id, uint32_t stride, EBufferFlag flag) -> void * {
        desc.name = id, desc.stride_in_byte = stride, desc.flag = flag;
        Buffer *b = buf_mngr->AllocBuffer(desc);
        return b ? b->cuda_ptr : nullptr;
    };
    m_params.accum_buffer = alloc("pt accum buffer", sizeof(float) * 4, EBufferFlag::None);
    m_params.albedo_buffer = alloc("albedo", sizeof(float) * 3, EBufferFlag::AllowDisplay);
    m_params.normal_buffer = alloc("normal", sizeof(float) * 3, EBufferFlag::AllowDisplay);
    m_params.test = alloc("test", sizeof(float), EBufferFlag::AllowDisplay);

    // the SBT of the reference (two hit records per render object, :172-237) has no equivalent: materials,
    // geometry and emitter offsets travel with the instances World hands to the back end
    m_params.handle = world->GetSceneHandle();
    m_dirty = true;
}

void PTPass::BindingEventCallback() noexcept {
    static bool bound = false;
    if (bound) return;
    bound = true;
    EventBinder<EWorldEvent::CameraChange>([](void *) {
        if (g_active_pass) g_active_pass->m_dirty = true;
    });
    EventBinder<EWorldEvent::RenderInstanceUpdate>([](void *) {
        if (g_active_pass) g_active_pass->m_dirty = true;
    });
    EventBinder<ESystemEvent::SceneLoad>([](void *p) {
        if (g_active_pass) g_active_pass->SetScene(static_cast<world::World *>(p));
    });
}

void PTPass::Inspector() noexcept {
    Pass::Inspector();
    Log::Info("sample count: %u, max trace depth: %d, accumulate: %d", m_params.sample_cnt + 1, m_max_depth, (int)m_accumulated_flag);
}
void PTPass::SetMaxDepth(int max_depth) noexcept {
    m_max_depth = std::clamp(max_depth, 1, 128);
    if ((int)m_params.config.max_depth != m_max_depth) m_dirty = true;
}
void PTPass::SetAccumulate(bool accumulate) noexcept {
    if (m_accumulated_flag != accumulate) m_dirty = true;
    m_accumulated_flag = accumulate;
}
void PTPass::Restart(unsigned int first_seed, unsigned int seed_stride) noexcept {
    m_first_seed = first_seed, m_seed_stride = seed_stride ? seed_stride : 1;
    m_dirty = true;
}
namespace {
struct CheckpointHeader { // little-endian, 64 bytes
    char magic[8];        // "PB2CKPT1"
    uint32_t width, height, max_depth, accumulate, sum_mode;
    uint32_t sample_cnt, random_seed, first_seed, seed_stride, frames_per_run;
    uint32_t scene_hash[2]; // SceneFingerprint(): 0, 0 in files written before the fingerprint existed (accepted)
    uint32_t reserved[2];
};
static_assert(sizeof(CheckpointHeader) == 64);
constexpr char kCheckpointMagic[8] = { 'P', 'B', '2', 'C', 'K', 'P', 'T', '1' };
// FNV-1a over what decides the image besides the seeds: camera matrices, every render object's transform, geometry size,
// flags and material (values, not device handles), the emitter table's weights and the integrator depth is in the header
// already.  A checkpoint of another scene or of other camera / instance edits at the same resolution is refused.
uint64_t SceneFingerprint(world::World *w) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    };
    const util::Mat4 s2c = w->camera->GetSampleToCameraMatrix(), c2w = w->camera->GetToWorldMatrix();
    mix(s2c.e, sizeof s2c.e), mix(c2w.e, sizeof c2w.e);
    for (world::RenderObject *ro : w->GetRenderobjects()) {
        mix(ro->transform.matrix.e, sizeof ro->transform.matrix.e);
        const uint32_t geo[4] = { (uint32_t)ro->geo_type, ro->sub_emitters_num, (uint32_t)ro->flip_normals | (uint32_t)ro->flip_tex_coords << 1 | (uint32_t)ro->is_emitter << 2,
                                  ro->shape && ro->geo_type == world::RenderObject::EGeoType::TriMesh ? ro->shape->mesh.vertex_num : 0u };
        mix(geo, sizeof geo);
        const pb2_material &m = ro->mat;
        mix(&m.type, 24); // type, twosided, eta, nonlinear, int_fdr, specular_sampling_weight
        for (const pb2_texture &t : m.tex) mix(&t, offsetof(pb2_texture, pad0)); // type, a, b, r0, r1 — not the bitmap handle
    }
    for (const pb2_emitter &e : w->emitters->GetAreaEmitters()) mix(&e.type, 12), mix(&e.area, sizeof e.area);
    if (const pb2_emitter *env = w->emitters->GetEnvEmitter()) mix(&env->type, 12), mix(env->radiance.a, sizeof env->radiance.a);
    return h;
}
}// namespace

bool PTPass::SaveCheckpoint(const std::filesystem::path &file) noexcept {
    try {
        if (!m_world || !m_params.accum_buffer || !m_params.frame_buffer) {
            Log::Warn("checkpoint: no scene to save");
            return false;
        }
        Synchronize();
        CheckpointHeader h{};
        std::memcpy(h.magic, kCheckpointMagic, 8);
        h.width = m_params.config.frame.width, h.height = m_params.config.frame.height;
        // a pending reset (m_dirty) has not reached m_params yet: a checkpoint taken then describes a pass that starts over
        h.max_depth = m_dirty ? static_cast<uint32_t>(m_max_depth) : m_params.config.max_depth;
        h.accumulate = (m_dirty ? m_accumulated_flag : m_params.config.accumulated_flag) ? 1u : 0u;
        h.sum_mode = m_sum_mode ? 1u : 0u;
        h.sample_cnt = m_dirty ? 0u : m_params.sample_cnt, h.random_seed = m_dirty ? m_first_seed : m_params.random_seed;
        h.first_seed = m_first_seed, h.seed_stride = m_seed_stride, h.frames_per_run = m_frames_per_run;
        const uint64_t fp = SceneFingerprint(m_world);
        h.scene_hash[0] = (uint32_t)fp, h.scene_hash[1] = (uint32_t)(fp >> 32);
        std::vector<float> accum(m_output_pixel_num * 4), frame(m_output_pixel_num * 4);
        if (pb2_download(accum.data(), m_params.accum_buffer, accum.size() * sizeof(float)) != PB2_OK ||
            pb2_download(frame.data(), m_params.frame_buffer, frame.size() * sizeof(float)) != PB2_OK) {
            Log::Error("checkpoint: %s", pb2_last_error());
            return false;
        }
        // written under a temporary name and renamed: a run killed while saving leaves the previous checkpoint intact
        const std::filesystem::path tmp = file.string() + ".tmp";
        std::FILE *f = std::fopen(tmp.string().c_str(), "wb");
        if (!f) {
            Log::Warn("checkpoint: cannot write %s", tmp.string().c_str());
            return false;
        }
        bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(accum.data(), sizeof(float), accum.size(), f) == accum.size() &&
                  std::fwrite(frame.data(), sizeof(float), frame.size(), f) == frame.size();
        ok = std::fclose(f) == 0 && ok;
        std::error_code ec;
        if (ok) std::filesystem::rename(tmp, file, ec);
        if (!ok || ec) {
            Log::Warn("checkpoint: cannot write %s", file.string().c_str());
            std::filesystem::remove(tmp, ec);
            return false;
        }
        return true;
    } catch (...) {
        return false;
    }
}

bool PTPass::LoadCheckpoint(const std::filesystem::path &file) noexcept {
    try {
        if (!m_world || !m_params.accum_buffer || !m_params.frame_buffer) {
            Log::Warn("checkpoint: set the scene before loading a checkpoint");
            return false;
        }
        std::FILE *f = std::fopen(file.string().c_str(), "rb");
        if (!f) {
            Log::Warn("checkpoint: cannot read %s", file.string().c_str());
            return false;
        }
        CheckpointHeader h{};
        std::vector<float> accum(m_output_pixel_num * 4), frame(m_output_pixel_num * 4);
        bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, kCheckpointMagic, 8) == 0;
        if (ok && (h.width != m_params.config.frame.width || h.height != m_params.config.frame.height)) {
            Log::Warn("checkpoint: %s holds a %ux%u frame, the scene renders %ux%u", file.string().c_str(), h.width, h.height, m_params.config.frame.width,
                      m_params.config.frame.height);
            ok = false;
        }
        if (ok && (h.scene_hash[0] | h.scene_hash[1])) {
            const uint64_t fp = SceneFingerprint(m_world);
            if (h.scene_hash[0] != (uint32_t)fp || h.scene_hash[1] != (uint32_t)(fp >> 32)) {
                Log::Warn("checkpoint: %s was taken from another scene, camera or instance placement", file.string().c_str());
                ok = false;
            }
        }
        ok = ok && std::fread(accum.data(), sizeof(float), accum.size(), f) == accum.size() && std::fread(frame.data(), sizeof(float), frame.size(), f) == frame.size();
        ok = ok && std::fgetc(f) == EOF; // nothing may follow
        std::fclose(f);
        if (!ok) {
            Log::Warn("checkpoint: %s is not a checkpoint of this scene", file.string().c_str());
            return false;
        }
        Synchronize();
        if (pb2_upload(m_params.accum_buffer, accum.data(), accum.size() * sizeof(float)) != PB2_OK ||
            pb2_upload(m_params.frame_buffer, frame.data(), frame.size() * sizeof(float)) != PB2_OK) {
            Log::Error("checkpoint: %s", pb2_last_error());
            return false;
        }
        m_max_depth = std::clamp(static_cast<int>(h.max_depth), 1, 128), m_accumulated_flag = h.accumulate != 0, m_sum_mode = h.sum_mode != 0;
        m_first_seed = h.first_seed, m_seed_stride = h.seed_stride ? h.seed_stride : 1, m_frames_per_run = h.frames_per_run ? h.frames_per_run : 1;
        m_params.config.max_depth = m_max_depth, m_params.config.accumulated_flag = m_accumulated_flag;
        m_params.sample_cnt = h.sample_cnt, m_params.random_seed = h.random_seed;
        m_dirty = false; // the next OnRun continues instead of starting over
        return true;
    } catch (...) {
        return false;
    }
}

pb2_render_stats PTPass::GetRenderStats() noexcept {
    pb2_render_stats st{};
    if (m_params.handle) pb2_render_stats_get(m_params.handle, &st);
    return st;
}
}// namespace Pupil::pt
