// Pupil::pt::PTPass — the path-tracing pass of example/path_tracer (pt_pass.{h,cpp}, type.h), same class and
// buffer names, with optix::Pass::Run + Synchronize replaced by pb2_render + pb2_synchronize.
//
// One OnRun() is one frame = one sample per pixel, exactly as in the reference (pt_pass.cpp:39-57); the
// progressive state is (accum buffer, sample_cnt, random_seed), reset whenever the pass is dirty.
// SetFramesPerRun(n) lets one OnRun() execute n consecutive frames inside the back end (same seeds, same
// running mean, no host round trip in between) — n = 1 is the reference behaviour.
#pragma once
#include "system.h"

#include <atomic>
#include <filesystem>

namespace Pupil::pt {
// pt::OptixLaunchParams (type.h:9-33): camera, emitter group and the AS handle live in the pb2 scene
struct LaunchParams {
    struct {
        unsigned int max_depth = 1;
        bool accumulated_flag = true;
        struct {
            unsigned int width = 0, height = 0;
        } frame;
    } config;
    unsigned int random_seed = 0;
    unsigned int sample_cnt = 0;
    void *accum_buffer = nullptr, *frame_buffer = nullptr, *normal_buffer = nullptr, *albedo_buffer = nullptr, *test = nullptr;
    pb2_scene *handle = nullptr;
};

class PTPass : public Pass {
public:
    PTPass(std::string_view name = "Path Tracing") noexcept;
    ~PTPass() noexcept override;
    void OnRun() noexcept override;
    void Inspector() noexcept override;
    void SetScene(world::World *world) noexcept;

    // the two inspector knobs of the reference (pt_pass.cpp:256-268)
    void SetMaxDepth(int max_depth) noexcept;
    void SetAccumulate(bool accumulate) noexcept;
    void SetFramesPerRun(unsigned int n) noexcept { m_frames_per_run = n ? n : 1; }
    // restart the progressive sequence at a given seed without touching the scene (checkpoint / shard support)
    void Restart(unsigned int first_seed = 0, unsigned int seed_stride = 1) noexcept;
    // Checkpoint of the progressive state — the reference has none; its state is (accum buffer, sample_cnt,
    // random_seed) (pt_pass.cpp:55-56), which is what the file holds next to the frame size and the pass settings.  A pass
    // resumed from a checkpoint of the same scene continues with the same seeds and the same running mean: the images are
    // bit-identical to an uninterrupted run.  Load wants the scene of the checkpoint set first (frame size is checked).
    bool SaveCheckpoint(const std::filesystem::path &file) noexcept;
    bool LoadCheckpoint(const std::filesystem::path &file) noexcept;
    void SetSumMode(bool sum) noexcept { m_sum_mode = sum, m_dirty = true; } // accumulate plain sums (multi-GPU shards)
    // Multi-GPU: this process is rank `rank` of `world`, one process per GPU.  Each OnRun renders this rank's
    // share of the step's seeds (pb2_shard_plan; strong: frames_per_run is the step's TOTAL sample count, split over the
    // ranks; weak: every rank renders frames_per_run) into plain sums and queues pb2_comm_reduce_frames, which leaves
    // sum / total spp in "final result".  OnRun does not wait: the reduction overlaps the next OnRun's render up to its first
    // accumulate kernel.  Synchronize() waits for both.  comm == nullptr switches sharding off.
    void SetShard(pb2_comm *comm, int rank, int world, bool strong, int reduce_mode) noexcept;
    void Synchronize() noexcept;
    bool IsSharded() const noexcept { return m_comm != nullptr; }
    const LaunchParams &GetLaunchParams() const noexcept { return m_params; }
    pb2_render_stats GetRenderStats() noexcept;

private:
    void BindingEventCallback() noexcept;
    void OnRunSharded() noexcept;
    LaunchParams m_params;
    size_t m_output_pixel_num = 0;
    std::atomic_bool m_dirty = true;
    world::World *m_world = nullptr;
    int m_max_depth = 1;
    bool m_accumulated_flag = true, m_sum_mode = false;
    unsigned int m_frames_per_run = 1, m_first_seed = 0, m_seed_stride = 1;
    pb2_comm *m_comm = nullptr; // not owned
    int m_rank = 0, m_n_ranks = 1, m_reduce_mode = PB2_REDUCE_ROOT;
    bool m_strong = false;
    unsigned int m_shard_step = 0, m_shard_total_spp = 0; // OnRun calls and samples of all ranks since the last restart
};
}// namespace Pupil::pt
