// Pupil::System / Pass / BufferManager — the render-pass interface of the reference, run headless.
//
//   Pass, EPassTag            framework/system/pass.{h,cpp}
//   Buffer, BufferManager     framework/system/buffer.{h,cpp}   (named device buffers, zero-initialised)
//   System, ESystemEvent      framework/system/system.{h,cpp}
//
// Differences, all forced by "headless on Linux": no GuiPass / DX12 shared buffers; System::Run() drives the
// passes on the calling thread and returns after `max_frames` frames (the reference loops until the window
// closes); cuda::Context::Init + optix::Context::Init become pb2_init(device).
#pragma once
#include "util.h"
#include "world.h"

#include <filesystem>
#include <memory>
#include <vector>

namespace Pupil {
enum class EPassTag : uint32_t { None = 0, Pre = 1 << 0, Post = 1 << 1, Asyn = 1 << 2 };
inline bool operator&(EPassTag target, EPassTag tag) noexcept { return static_cast<uint32_t>(target) & static_cast<uint32_t>(tag); }
inline EPassTag operator|(EPassTag a, EPassTag b) noexcept { return static_cast<EPassTag>(static_cast<uint32_t>(a) | static_cast<uint32_t>(b)); }

class Pass {
protected:
    Timer m_timer;
    double m_last_exec_time = 0.;
    bool m_enable = true;

public:
    const std::string name;
    const EPassTag tag;
    Pass(std::string_view name, EPassTag tag = EPassTag::None) noexcept : name(name), tag(tag) {}
    virtual ~Pass() = default;
    virtual void Run() noexcept; // timed wrapper around OnRun (pass.cpp:6-13)
    virtual void Inspector() noexcept;
    virtual void OnRun() noexcept = 0;
    void Toggle() noexcept { m_enable ^= true; }
    void SetEnablility(bool enable) noexcept { m_enable = enable; }
    bool IsEnabled() const noexcept { return m_enable; }
    double LastExecTimeMs() const noexcept { return m_last_exec_time; }
};

enum class EBufferFlag : unsigned int { None = 0, SharedWithDX12 = 1, AllowDisplay = 1 << 1 };
struct BufferDesc {
    const char *name = nullptr;
    EBufferFlag flag = EBufferFlag::None;
    uint32_t width = 1, height = 1, stride_in_byte = 1;
};
struct Buffer {
    BufferDesc desc{};
    std::string name;
    void *cuda_ptr = nullptr; // CUdeviceptr in the reference
    Buffer() noexcept = default;
    explicit Buffer(const BufferDesc &d) noexcept : desc(d), name(d.name ? d.name : "") { desc.name = name.c_str(); }
    ~Buffer() noexcept;
    size_t SizeInBytes() const noexcept { return static_cast<size_t>(desc.width) * desc.height * desc.stride_in_byte; }
};
class BufferManager : public util::Singleton<BufferManager> {
public:
    constexpr static std::string_view DEFAULT_FINAL_RESULT_BUFFER_NAME = "final result";
    void Destroy() noexcept;
    [[nodiscard]] Buffer *GetBuffer(std::string_view id) noexcept;
    Buffer *AllocBuffer(const BufferDesc &desc) noexcept; // replaces an existing buffer of the same name
    [[nodiscard]] const std::vector<std::string> &GetBufferNameList() const noexcept { return m_buffer_names; }

private:
    std::unordered_map<std::string, std::unique_ptr<Buffer>> m_buffers;
    std::vector<std::string> m_buffer_names;
};

enum class ESystemEvent { Quit, Precompute, StartRendering, StopRendering, SceneLoad, FrameFinished };

class System : public util::Singleton<System> {
public:
    bool render_flag = true;
    bool quit_flag = false;
    uint64_t max_frames = 0; // headless: Run() returns after this many frames (0 = until quit_flag)
    int device = 0;          // CUDA device used by Init

    void Init(bool has_window = true) noexcept;
    void Run() noexcept;
    void Destroy() noexcept;
    void AddPass(Pass *pass) noexcept;
    void RemovePass(Pass *pass) noexcept;
    // The reference's SetScene returns nothing and carries on after a failed load (system.cpp:143-165); here the outcome is
    // returned so that the C entry points can fail, and a failed load leaves NO scene behind (passes are told with a null
    // SceneLoad event, the device scene is cleared) instead of the previous scene's buffers and geometry.
    bool SetScene(std::filesystem::path scene_file_path) noexcept;
    // same hand-off for a scene assembled in memory (SceneDesc route): world->scene must already be filled
    bool SetScene(resource::Scene *scene) noexcept;
    bool IsInitialized() const noexcept { return m_initialized; }
    uint64_t FramesRendered() const noexcept { return m_frames; }

private:
    void AfterSceneLoad() noexcept;
    void AfterSceneLoadFailed() noexcept;
    std::vector<Pass *> m_passes, m_pre_passes;
    Timer m_render_timer;
    bool m_initialized = false, m_scene_load_flag = false, m_system_run_flag = false;
    uint64_t m_frames = 0;
};
}// namespace Pupil
