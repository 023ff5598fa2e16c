// path_tracer — headless counterpart of example/path_tracer/main.cpp:5-22:
//   System::Init -> AddPass(PTPass) -> SetScene(xml) -> Run -> Destroy
// plus what a window-less run needs: an spp limit and an image file.
//   path_tracer --scene file.xml [--spp 64] [--depth N] [--device 0] [--out image.pfm|.exr|.hdr|.png] [--batch 16] [--builder 0|1]
//               [--checkpoint file] [--resume file]
// --spp is the total sample count.  --checkpoint writes the progressive state (PTPass::SaveCheckpoint) after every batch, so a
// killed run loses at most one batch; --resume continues from such a file up to --spp (bit-identical to an uninterrupted run).
#include "pt_pass.h"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include "image.h"
#include <filesystem>
#include <memory>
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
    int depth = 0, device = 0, builder = -1;
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
        else if (!std::strcmp(argv[i], "--verbose")) Log::level = 2;
        else {
            std::fprintf(stderr, "usage: %s --scene file.xml [--spp N] [--depth N] [--device D] [--out image.pfm|.exr|.hdr|.png] [--batch N] [--builder 0|1] [--checkpoint file] [--resume file]\n", argv[0]);
            return 2;
        }
    }
    if (!scene_path) {
        std::fprintf(stderr, "path_tracer: --scene is required\n");
        return 2;
    }
    auto *system = util::Singleton<System>::instance();
    system->device = device;
    system->Init(false);
    if (!system->IsInitialized()) return 1; // no CUDA device: there is no CPU path
    int rc = 0;
    {
        auto pt_pass = std::make_unique<pt::PTPass>("Path Tracing");
        system->AddPass(pt_pass.get());
        if (builder >= 0) util::Singleton<world::World>::instance()->SetBvhBuilder(builder);
        system->SetScene(std::filesystem::path(scene_path));
        if (!pt_pass->GetLaunchParams().accum_buffer) {
            std::fprintf(stderr, "path_tracer: could not load %s\n", scene_path);
            rc = 1;
        } else {
            if (depth > 0) pt_pass->SetMaxDepth(depth);
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
            timer.Stop();
            spp = std::max(done, first) - first; // samples rendered by this run, for the rate below
            const auto &lp = pt_pass->GetLaunchParams();
            const uint32_t w = lp.config.frame.width, h = lp.config.frame.height;
            std::vector<float> img(static_cast<size_t>(w) * h * 4);
            pb2_download(img.data(), lp.frame_buffer, img.size() * sizeof(float));
            const auto &bs = util::Singleton<world::World>::instance()->GetBuildStats();
            std::printf("%ux%u, %u spp (%u in total), depth %u: %.1f ms (%.2f Msamples/s); BVH: %llu prims, %llu nodes, %.2f ms\n", w, h, spp, done, lp.config.max_depth,
                        timer.ElapsedMilliseconds(), 1e-3 * w * h * spp / timer.ElapsedMilliseconds(), (unsigned long long)bs.n_prims,
                        (unsigned long long)bs.n_nodes, bs.build_ms);
            if (!WriteImage(out_path, img, w, h)) std::fprintf(stderr, "path_tracer: cannot write %s\n", out_path), rc = 1;
        }
        system->RemovePass(pt_pass.get());
    }
    system->Destroy();
    return rc;
}
