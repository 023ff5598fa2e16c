// path_tracer — headless counterpart of example/path_tracer/main.cpp:5-22:
//   System::Init -> AddPass(PTPass) -> SetScene(xml) -> Run -> Destroy
// plus what a window-less run needs: an spp limit and an image file.
//   path_tracer --scene file.xml [--spp 64] [--depth N] [--device 0] [--out image.pfm|.exr|.hdr|.png] [--batch 16] [--builder 0|1]
//               [--checkpoint file] [--resume file] [--gpus N]
// --gpus N renders on N GPUs of this box, one process per GPU: the process started by the user is rank 0 on --device, it
// starts ranks 1..N-1 (this executable again, on devices --device + rank) and hands them the NCCL id on the command line.
// Every rank loads the scene and renders its share of each batch's seeds (PTPass::SetShard: seed = base + rank + k N, plain
// sums); pb2_comm_reduce_frames combines them over NVLink and rank 0 writes the image.
// --spp is the total sample count.  --checkpoint writes the progressive state (PTPass::SaveCheckpoint) after every batch, so a
// killed run loses at most one batch; --resume continues from such a file up to --spp (bit-identical to an uninterrupted run).
#include "pt_pass.h"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include "image.h"
#include <filesystem>
#include <memory>
#include <string>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

using namespace Pupil;

// The reference's screenshot path (util::BitmapTexture::Save, util/texture.cpp:13-85: hdr or exr) plus PFM; the format
// follows the file extension.
static bool WriteImage(const char *path, const std::vector<float> &rgba, uint32_t w, uint32_t h) {
    const std::string ext = std::filesystem::path(path).extension().string();
    const util::EImageFileFormat fmt = ext == ".exr" ? util::EImageFileFormat::EXR : ext == ".hdr" ? util::EImageFileFormat::HDR : ext == ".png" ? util::EImageFileFormat::PNG : util::EImageFileFormat::PFM;
    return util::SaveImage(rgba.data(), w, h, path, fmt);
}

int main(int argc, char **argv) {
    const char *scene_path = nullptr, *out_path = "path_tracer.pfm", *checkpoint_path = nullptr, *resume_path = nullptr;
    unsigned spp = 64, batch = 16;
    int depth = 0, device = 0, builder = -1, gpus = 1, rank = 0;
    const char *comm_id_hex = nullptr;
    for (int i = 1; i < argc; ++i) {
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (!std::strcmp(argv[i], "--scene")) scene_path = next();
        else if (!std::strcmp(argv[i], "--spp")) spp = std::max(1, std::atoi(next()));
        else if (!std::strcmp(argv[i], "--depth")) depth = std::atoi(next());
        else if (!std::strcmp(argv[i], "--device")) device = std::atoi(next());
        else if (!std::strcmp(argv[i], "--out")) out_path = next();
        else if (!std::strcmp(argv[i], "--batch")) batch = std::max(1, std::atoi(next()));
        else if (!std::strcmp(argv[i], "--builder")) builder = std::atoi(next());
        else if (!std::strcmp(argv[i], "--checkpoint")) checkpoint_path = next();
        else if (!std::strcmp(argv[i], "--resume")) resume_path = next();
        else if (!std::strcmp(argv[i], "--gpus")) gpus = std::max(1, std::atoi(next()));
        else if (!std::strcmp(argv[i], "--rank")) rank = std::atoi(next());          // set by rank 0 for the processes it starts
        else if (!std::strcmp(argv[i], "--comm-id")) comm_id_hex = next();           // likewise: the NCCL id, 256 hex digits
        else if (!std::strcmp(argv[i], "--verbose")) Log::level = 2;
        else {
            std::fprintf(stderr, "usage: %s --scene file.xml [--spp N] [--depth N] [--device D] [--out image.pfm|.exr|.hdr|.png] [--batch N] [--builder 0|1] [--checkpoint file] [--resume file] [--gpus N]\n", argv[0]);
            return 2;
        }
    }
    if (!scene_path) {
        std::fprintf(stderr, "path_tracer: --scene is required\n");
        return 2;
    }
    // ---- multi-GPU bootstrap: rank 0 creates the NCCL id and starts the other ranks ----
    uint8_t comm_id[PB2_COMM_ID_BYTES] = {};
    std::vector<pid_t> children;
    if (gpus > 1 && (checkpoint_path || resume_path)) {
        std::fprintf(stderr, "path_tracer: --checkpoint / --resume hold one GPU's running mean; they cannot be combined with --gpus\n");
        return 2;
    }
    if (gpus > 1 && rank == 0) {
        if (pb2_comm_unique_id(comm_id) != PB2_OK) {
            std::fprintf(stderr, "path_tracer: %s\n", pb2_last_error());
            return 1;
        }
        std::string hex;
        for (uint8_t b : comm_id) {
            char t[3];
            std::snprintf(t, sizeof t, "%02x", b);
            hex += t;
        }
        for (int r = 1; r < gpus; ++r) {
            const pid_t pid = fork();
            if (pid == 0) {
                std::vector<std::string> args(argv, argv + argc);
                args.insert(args.end(), { "--rank", std::to_string(r), "--comm-id", hex });
                std::vector<char *> cargs;
                for (auto &a : args) cargs.push_back(a.data());
                cargs.push_back(nullptr);
                execv("/proc/self/exe", cargs.data());
                std::perror("path_tracer: execv");
                _exit(127);
            }
            if (pid < 0) {
                std::perror("path_tracer: fork");
                return 1;
            }
            children.push_back(pid);
        }
    } else if (gpus > 1) {
        if (!comm_id_hex || std::strlen(comm_id_hex) != 2 * PB2_COMM_ID_BYTES || rank < 0 || rank >= gpus) {
            std::fprintf(stderr, "path_tracer: --rank / --comm-id are set by rank 0\n");
            return 2;
        }
        for (int k = 0; k < PB2_COMM_ID_BYTES; ++k) {
            const char t[3] = { comm_id_hex[2 * k], comm_id_hex[2 * k + 1], 0 };
            comm_id[k] = static_cast<uint8_t>(std::strtoul(t, nullptr, 16));
        }
    }
    auto wait_children = [&]() {
        int bad = 0;
        for (pid_t pid : children) {
            int status = 0;
            if (waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) ++bad;
        }
        children.clear();
        return bad;
    };
    auto *system = util::Singleton<System>::instance();
    system->device = device + rank;
    system->Init(false);
    if (!system->IsInitialized()) { // no CUDA device: there is no CPU path
        wait_children();
        return 1;
    }
    int rc = 0;
    pb2_comm *comm = nullptr;
    {
        auto pt_pass = std::make_unique<pt::PTPass>("Path Tracing");
        system->AddPass(pt_pass.get());
        if (builder >= 0) util::Singleton<world::World>::instance()->SetBvhBuilder(builder);
        if (!system->SetScene(std::filesystem::path(scene_path)) || !pt_pass->GetLaunchParams().accum_buffer) {
            std::fprintf(stderr, "path_tracer: could not load %s\n", scene_path);
            rc = 1;
        } else {
            if (depth > 0) pt_pass->SetMaxDepth(depth);
            if (gpus > 1) {
                if (pb2_comm_create(&comm, gpus, rank, comm_id) != PB2_OK) {
                    std::fprintf(stderr, "path_tracer: rank %d: %s\n", rank, pb2_last_error());
                    rc = 1;
                } else {
                    pt_pass->SetShard(comm, rank, gpus, /*strong=*/true, PB2_REDUCE_ROOT); // --spp is the total over all GPUs; rank 0 writes the file
                }
            }
            unsigned done = 0;
            if (resume_path) {
                if (!pt_pass->LoadCheckpoint(resume_path)) {
                    std::fprintf(stderr, "path_tracer: cannot resume from %s\n", resume_path);
                    rc = 1;
                } else {
                    done = pt_pass->GetLaunchParams().sample_cnt;
                    std::printf("resumed at %u spp\n", done);
                }
            }
            const unsigned first = done;
            Timer timer;
            timer.Start();
            while (rc == 0 && done < spp) { // whole batches, then the remainder
                const unsigned n = std::min(batch, spp - done);
                pt_pass->SetFramesPerRun(n);
                system->max_frames = 1;
                system->Run();
                done += n;
                if (checkpoint_path && !pt_pass->SaveCheckpoint(checkpoint_path)) std::fprintf(stderr, "path_tracer: cannot write %s\n", checkpoint_path), rc = 1;
            }
            pt_pass->Synchronize();
            timer.Stop();
            spp = std::max(done, first) - first; // samples rendered by this run, for the rate below
            const auto &lp = pt_pass->GetLaunchParams();
            const uint32_t w = lp.config.frame.width, h = lp.config.frame.height;
            std::vector<float> img(static_cast<size_t>(w) * h * 4);
            if (rank == 0 && rc == 0) pb2_download(img.data(), lp.frame_buffer, img.size() * sizeof(float));
            const auto &bs = util::Singleton<world::World>::instance()->GetBuildStats();
            if (rank == 0 && rc == 0) {
                std::printf("%ux%u, %u spp (%u in total) on %d GPU%s, depth %u: %.1f ms (%.2f Msamples/s); BVH: %llu prims, %llu nodes, %.2f ms\n", w, h, spp, done, gpus,
                            gpus > 1 ? "s" : "", lp.config.max_depth, timer.ElapsedMilliseconds(), 1e-3 * w * h * spp / timer.ElapsedMilliseconds(),
                            (unsigned long long)bs.n_prims, (unsigned long long)bs.n_nodes, bs.build_ms);
                if (!WriteImage(out_path, img, w, h)) std::fprintf(stderr, "path_tracer: cannot write %s\n", out_path), rc = 1;
            }
        }
        pt_pass->SetShard(nullptr, 0, 1, false, PB2_REDUCE_ROOT);
        if (comm) pb2_comm_destroy(comm);
        system->RemovePass(pt_pass.get());
    }
    system->Destroy();
    if (wait_children() != 0) std::fprintf(stderr, "path_tracer: a rank failed\n"), rc = 1;
    return rc;
}
