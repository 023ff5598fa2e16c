#include "pt_pass.h"

#include <algorithm>

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
        if (m_sum_mode) pb2_memset(m_params.accum_buffer, 0, m_output_pixel_num * sizeof(float) * 4); // sums start from zero
        m_dirty = false;
    }
    m_params.handle = m_world->GetSceneHandle(); // refreshes BVH / camera / emitters when they changed
    if (!m_params.handle) return;

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

void PTPass::SetScene(world::World *world) noexcept {
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
pb2_render_stats PTPass::GetRenderStats() noexcept {
    pb2_render_stats st{};
    if (m_params.handle) pb2_render_stats_get(m_params.handle, &st);
    return st;
}
}// namespace Pupil::pt
