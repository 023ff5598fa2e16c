#include "system.h"

#include <algorithm>
#include <mutex>

namespace Pupil {
namespace {
std::mutex m_render_system_mutex;
}

void Pass::Run() noexcept {
    if (!m_enable) return;
    m_timer.Start();
    OnRun();
    m_timer.Stop();
    m_last_exec_time = m_timer.ElapsedMilliseconds();
}
void Pass::Inspector() noexcept { Log::Info("pass [%s]: %.3f ms", name.c_str(), m_last_exec_time); }

Buffer::~Buffer() noexcept {
    if (cuda_ptr) pb2_free(cuda_ptr);
}
void BufferManager::Destroy() noexcept { m_buffers.clear(), m_buffer_names.clear(); }
Buffer *BufferManager::GetBuffer(std::string_view id) noexcept {
    auto it = m_buffers.find(std::string(id));
    return it == m_buffers.end() ? nullptr : it->second.get();
}
Buffer *BufferManager::AllocBuffer(const BufferDesc &desc) noexcept {
    auto buffer = std::make_unique<Buffer>(desc);
    if (pb2_malloc(&buffer->cuda_ptr, buffer->SizeInBytes()) != PB2_OK) { // zero-filled, like buffer.cpp:44-45
        Log::Error("buffer [%s]: %s", buffer->name.c_str(), pb2_last_error());
        return nullptr;
    }
    Buffer *raw = buffer.get();
    if (m_buffers.find(raw->name) == m_buffers.end()) m_buffer_names.push_back(raw->name);
    m_buffers[raw->name] = std::move(buffer);
    return raw;
}

void System::Init(bool has_window) noexcept {
    if (has_window) Log::Warn("the DX12 / ImGui window is Windows-only: running headless");
    if (pb2_init(device) != PB2_OK) { // cuda::Context::Init + optix::Context::Init
        Log::Error("pb2_init(%d): %s", device, pb2_last_error());
        return;
    }
    util::Singleton<world::World>::instance()->Init();
    static bool bound = false;
    if (!bound) {
        bound = true;
        EventBinder<ESystemEvent::Quit>([this](void *) { quit_flag = true; });
        EventBinder<ESystemEvent::StartRendering>([this](void *) { render_flag = true, m_scene_load_flag = true; });
        EventBinder<ESystemEvent::StopRendering>([this](void *) { render_flag = false; });
        EventBinder<ESystemEvent::Precompute>([this](void *) {
            for (auto *pass : m_pre_passes) pass->Run();
        });
    }
    quit_flag = false, m_frames = 0;
    m_initialized = true;
}

void System::Run() noexcept {
    m_system_run_flag = true;
    if (m_scene_load_flag) {
        EventDispatcher<ESystemEvent::Precompute>();
        EventDispatcher<ESystemEvent::StartRendering>();
    } else {
        EventDispatcher<ESystemEvent::StopRendering>();
    }
    const uint64_t stop_at = max_frames ? m_frames + max_frames : 0;
    while (!quit_flag && render_flag && (stop_at == 0 || m_frames < stop_at)) {
        std::unique_lock render_lock(m_render_system_mutex);
        m_render_timer.Start();
        for (auto *pass : m_passes) pass->Run();
        m_render_timer.Stop();
        ++m_frames;
        EventDispatcher<ESystemEvent::FrameFinished>(m_render_timer.ElapsedMilliseconds());
    }
    m_system_run_flag = false;
}

void System::Destroy() noexcept {
    util::Singleton<BufferManager>::instance()->Destroy();
    util::Singleton<world::World>::instance()->Destroy();
    m_passes.clear(), m_pre_passes.clear();
    m_initialized = false, m_scene_load_flag = false;
}

void System::AddPass(Pass *pass) noexcept {
    if (pass->tag & EPassTag::Pre) m_pre_passes.push_back(pass);
    else m_passes.push_back(pass);
}
void System::RemovePass(Pass *pass) noexcept {
    m_passes.erase(std::remove(m_passes.begin(), m_passes.end(), pass), m_passes.end());
    m_pre_passes.erase(std::remove(m_pre_passes.begin(), m_pre_passes.end(), pass), m_pre_passes.end());
}

void System::AfterSceneLoad() noexcept {
    auto *world = util::Singleton<world::World>::instance();
    auto *buf_mngr = util::Singleton<BufferManager>::instance();
    BufferDesc desc{};
    desc.name = buf_mngr->DEFAULT_FINAL_RESULT_BUFFER_NAME.data();
    desc.width = static_cast<uint32_t>(world->scene->sensor.film.w), desc.height = static_cast<uint32_t>(world->scene->sensor.film.h);
    desc.stride_in_byte = sizeof(float) * 4;
    buf_mngr->AllocBuffer(desc);
    m_scene_load_flag = true;
    EventDispatcher<ESystemEvent::SceneLoad>(static_cast<void *>(world));
}
void System::AfterSceneLoadFailed() noexcept {
    util::Singleton<world::World>::instance()->Reset();
    m_scene_load_flag = false;
    EventDispatcher<ESystemEvent::SceneLoad>(nullptr); // passes drop the previous scene's buffers and handle
}
bool System::SetScene(std::filesystem::path scene_file_path) noexcept {
    if (!std::filesystem::exists(scene_file_path)) {
        Log::Warn("scene file [%s] does not exist", scene_file_path.string().c_str());
        return false;
    }
    {
        std::unique_lock render_lock(m_render_system_mutex);
        if (!util::Singleton<world::World>::instance()->LoadScene(scene_file_path)) {
            Log::Warn("scene load failed");
            AfterSceneLoadFailed();
            return false;
        }
        AfterSceneLoad();
    }
    render_flag = true;
    if (m_system_run_flag) {
        EventDispatcher<ESystemEvent::Precompute>();
        EventDispatcher<ESystemEvent::StartRendering>();
    }
    return true;
}
bool System::SetScene(resource::Scene *scene) noexcept {
    std::unique_lock render_lock(m_render_system_mutex);
    if (!util::Singleton<world::World>::instance()->LoadScene(scene)) {
        Log::Warn("scene load failed");
        AfterSceneLoadFailed();
        return false;
    }
    EventDispatcher<EWorldEvent::CameraChange>();
    AfterSceneLoad();
    render_flag = true;
    return true;
}
}// namespace Pupil
