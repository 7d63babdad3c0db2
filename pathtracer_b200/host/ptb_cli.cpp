// ptb_cli.cpp — headless driver: the `rayTracer <scene> <out>` path of the reference (mainApp.cpp:38-49) over the
// CUDA library: a .scn file written by Raytracer::save_scene, or the synthetic scenes of SURVEY.md §8d.
//   ptb_cli [--gpus N] [--preset NAME] <scene.scn> <out.ppm> [W H spp]          (0 keeps the file's value)
//   ptb_cli [--gpus N] [--preset NAME] <C1|torus>  <out.ppm> [W H spp nv]
// --gpus N: devices 0..N-1 of this process render the frame together (tile-sharded, NCCL gather inside the library);
// --preset NAME: a material preset of the reference's object menu on the synthetic scene's objects (gold, gold_ngan, ..., copper_ngan).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ptb_raytracer.hpp"

using namespace ptbhost;

static std::shared_ptr<TriMesh> displaced_torus(int nv) {   // SURVEY.md §8d generator (same as pathtracer_b200/scenes.py)
    auto g = std::make_shared<TriMesh>();
    const int nu = 2 * nv;
    for (int j = 0; j <= nv; j++)
        for (int i = 0; i <= nu; i++) {
            const double u = 2 * M_PI * i / nu, v = 2 * M_PI * j / nv;
            const double n[3] = {std::cos(v) * std::cos(u), std::sin(v), std::cos(v) * std::sin(u)};
            const double rho = 0.4 * (1 + 0.15 * std::sin(9 * u) * std::sin(7 * v) + 0.04 * std::sin(40 * u + 3) * std::sin(33 * v));
            const double p[3] = {std::cos(u) + rho * n[0], rho * n[1], std::sin(u) + rho * n[2]};
            for (int k = 0; k < 3; k++) { g->vertices.push_back((float)p[k]); g->normals.push_back((float)n[k]); }
            g->uvs.push_back((float)((double)i / nu)); g->uvs.push_back((float)((double)j / nv));
        }
    for (int j = 0; j < nv; j++)
        for (int i = 0; i < nu; i++) {
            const int a = j * (nu + 1) + i, b = a + 1, c = a + nu + 1, d = c + 1;
            const int t[2][3] = {{a, b, c}, {b, d, c}};
            for (auto& tt : t) { for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) g->indices.push_back(tt[k]); g->indices.push_back(0); }
        }
    return g;
}

int main(int argc, char** argv) {
    int n_gpus = 1;
    const char* preset = nullptr;
    while (argc > 2 && argv[1][0] == '-' && argv[1][1] == '-') {
        if (!std::strcmp(argv[1], "--gpus")) n_gpus = std::atoi(argv[2]);
        else if (!std::strcmp(argv[1], "--preset")) preset = argv[2];
        else { std::fprintf(stderr, "ptb_cli: unknown option %s\n", argv[1]); return 2; }
        argv += 2; argc -= 2;
    }
    if (n_gpus < 1 || n_gpus > 64) { std::fprintf(stderr, "ptb_cli: --gpus must be 1..64\n"); return 2; }
    if (argc < 3) { std::fprintf(stderr, "usage: %s <scene.scn|curves.yarn|C1|torus> <out.ppm> [W H spp nv]   (PTB_FRAME=<n> renders frame n of a key-framed .scn)\n", argv[0]); return 2; }
    try {
        std::vector<int> devices;
        for (int i = 0; i < n_gpus; i++) devices.push_back(i);
        Raytracer rt(devices);
        const size_t len = std::strlen(argv[1]);
        const bool scn = len > 4 && !std::strcmp(argv[1] + len - 4, ".scn");
        if (scn) {
            if (const char* fr = std::getenv("PTB_FRAME")) rt.s.current_frame = std::atoi(fr);
            rt.load_scene(argv[1]);
            if (argc > 3 && std::atoi(argv[3]) > 0) rt.W = std::atoi(argv[3]);
            if (argc > 4 && std::atoi(argv[4]) > 0) rt.H = std::atoi(argv[4]);
            if (argc > 5 && std::atoi(argv[5]) > 0) rt.nrays = std::atoi(argv[5]);
        } else {
        rt.loadScene();
        rt.W = argc > 3 ? std::atoi(argv[3]) : 512; rt.H = argc > 4 ? std::atoi(argv[4]) : 512;
        rt.nrays = argc > 5 ? std::atoi(argv[5]) : 64; rt.nb_bounces = 5;
        auto phong = [](Vector kd, float ks, float ne) {
            Material m;
            m.set(PTB_SLOT_KD, Texture(kd)).set(PTB_SLOT_KS, Texture(ks)).set(PTB_SLOT_NE, Texture(ne)).set(PTB_SLOT_TRANSP, Texture(1.f)).set(PTB_SLOT_REFR, Texture(1.3f));
            return m;
        };
        if (len > 5 && !std::strcmp(argv[1] + len - 5, ".yarn")) {
            rt.s.addObject(std::make_shared<Yarns>(argv[1]));       // a dropped .yarn file: added as it is (mainApp.cpp:2413-2416)
        } else if (!std::strcmp(argv[1], "C1")) {
            auto s1 = std::make_shared<Sphere>(Vector(0, -17.3f, 0), 10.f); s1->materials.push_back(phong(Vector(.8f, .3f, .3f), 0.f, 1.f));
            auto s2 = std::make_shared<Sphere>(Vector(-15, -20.3f, 5), 7.f); s2->materials.push_back(phong(Vector(.3f, .8f, .3f), .3f, 50.f));
            rt.s.addObject(s1); rt.s.addObject(s2);
        } else {
            auto g = displaced_torus(argc > 6 ? std::atoi(argv[6]) : 100);
            g->scale = 30.f;
            g->max_translation = Vector(0, -27.3f + 0.29f * 30.f, 0);   // rests on the ground plane (mainApp.cpp:2404-2408)
            g->materials.push_back(phong(Vector(.5f, .5f, .5f), .2f, 50.f));
            rt.s.addObject(g);
        }
        if (preset) for (size_t i = 3; i < rt.s.objects.size(); i++) rt.s.objects[i]->set_preset(preset, 0);
        rt.commit();
        }
        rt.render_image_nopreviz();
        std::FILE* f = std::fopen(argv[2], "wb");
        if (!f) throw Error("cannot open output");
        std::fprintf(f, "P6\n%d %d\n255\n", rt.W, rt.H);
        std::fwrite(rt.image.data(), 1, rt.image.size(), f);
        std::fclose(f);
        const double rays = (double)rt.stats.rays_closest + (double)rt.stats.rays_shadow;
        std::printf("%d GPU(s), %dx%d %d spp: %.1f ms on the device, %.1f Msamples/s, %.1f Mrays/s, %llu kernel launches\n", n_gpus, rt.W, rt.H, rt.nrays, rt.stats.ms_device,
                    rt.stats.samples / rt.stats.ms_device / 1e3, rays / rt.stats.ms_device / 1e3, (unsigned long long)rt.stats.kernel_launches);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ptb_cli: %s\n", e.what());
        return 1;
    }
    return 0;
}
