// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Headless driver over the REFERENCE's own classes (Raytracer, Scene, TriMesh, Sphere, Plane,
// PhongBRDF, IsoMERLBRDF ...).  It is #included at the end of the unity translation unit that
// oracle/build_ref.py writes, after the reference's patched .cpp files, so every symbol used below
// is the reference's.  It exports the ABI of include/ptb200.h under the prefix ref_ so the parity
// tests can feed one scene description to the reference and to the CUDA path.
//
// Scene construction follows SURVEY.md App. B (in-memory route, no .scn/.obj files):
//   meshes : `new TriMesh()`, fill vertices/normals/uvs/indices, `init(scene,"synthetic",...)`
//            (TriangleMesh.cpp:718-841 skips the file readers for names without .obj/.off/.wrl)
//   slots  : Texture objects pushed into Object::{textures,specularmap,...} (Geometry.h:672)
//   envmap : Sphere::{envtex,envW,envH,has_envmap} filled directly (Geometry.h:1096-1101)
//   MERL   : IsoMERLBRDF with `data` pointed at the caller's table (BRDF.h:247)
#define ORACLE_PREFIX ref_
#include "prefix.h"
#include "ptb200.h"

#include <array>
#include <chrono>
#include <map>
#include <string>
#include <vector>

uint32_t ptb_ref_global_seed = 0;
unsigned long long ptb_ref_cnt[64][8];

struct ptb_ctx {
    Raytracer* rt;
    std::string err;
    std::vector<double*> merl;
    bool committed;
    int threads;
    double ms_build;
    long long n_tri;
    int frame = 0;
    std::map<const Object*, std::vector<int>> pts_perm;   // PointSet: position after build_bvh -> index handed in (the reference keeps no such map)
};

static std::string g_create_err;

static void apply_xform(Object* o, const ptb_xform* xf, bool is_mesh) {
    if (!xf) return;
    o->scale = xf->scale;
    for (int i = 0; i < 9; i++) o->mat_rotation[i] = xf->rotation[i];
    o->max_translation = Vector(xf->translation[0], xf->translation[1], xf->translation[2]);
    if (!(xf->rotation_center[0] != xf->rotation_center[0]))  // NaN keeps the object's own centre
        o->rotation_center = Vector(xf->rotation_center[0], xf->rotation_center[1], xf->rotation_center[2]);
}

static void apply_flags(Object* o, int flags) {
    o->miroir = (flags & PTB_OBJ_MIRROR) != 0;
    o->flip_normals = (flags & PTB_OBJ_FLIP_NORMALS) != 0;
    o->ghost = (flags & PTB_OBJ_GHOST) != 0;
}

static Texture make_tex(const ptb_tex& t, int type) {
    Texture tex;  // W=H=0, multiplier (1,1,1), type 0
    tex.type = type;
    tex.multiplier = Vector(t.mult[0], t.mult[1], t.mult[2]);
    tex.filename = "Null";
    if (t.texels && t.W > 0 && t.H > 0) {
        tex.W = t.W;
        tex.H = t.H;
        tex.values.assign(t.texels, t.texels + (size_t)t.W * t.H * 3);
    }
    return tex;
}

static void set_slot(std::vector<Texture>& slot, int group, const Texture& t) {
    if ((int)slot.size() <= group) slot.resize(group + 1);
    slot[group] = t;
}

extern "C" {

const char* ptb_version(void) { return "ptb-oracle-ref (reference sources, own BVH, headless)"; }

const char* ptb_last_error(const ptb_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int ptb_create(int device_id, ptb_ctx** out) {
    (void)device_id;
    if (!out) return PTB_ERR_INVALID;
    ptb_ctx* c = new ptb_ctx();
    c->rt = new Raytracer();  // heap: contribsArray is ~0.6 MB
    Raytracer* rt = c->rt;
    rt->W = 64; rt->H = 64; rt->nrays = 1; rt->nb_bounces = 5;
    rt->last_nrays = -1; rt->lastfilter = -1;
    rt->sigma_filter = 0.5f; rt->gamma = 2.2f;
    rt->autosave = false; rt->has_denoiser = false; rt->is_recording = false;
    rt->sphereEnv = NULL;
    Scene& s = rt->s;
    s.fog_density = s.fog_absorption = s.fog_density_decay = s.fog_absorption_decay = 0;
    s.fog_type = s.fog_phase_type = 0; s.phase_aniso = 0; s.nbframes = 1;
    s.lumiere = NULL; s.intensite_lumiere = 0; s.envmap_intensity = 1;
    c->committed = false;
    c->threads = 0;
    c->ms_build = 0;
    c->n_tri = 0;
    *out = c;
    return PTB_OK;
}

void ptb_destroy(ptb_ctx* c) {
    if (!c) return;
    c->rt->s.clear();
    delete c->rt;
    for (size_t i = 0; i < c->merl.size(); i++) free(c->merl[i]);
    delete c;
}

int ptb_add_sphere(ptb_ctx* c, const float O[3], float R, const ptb_xform* xf, int flags, int* out_id) {
    if (!c || !O) return PTB_ERR_INVALID;
    Sphere* sp = new Sphere(Vector(O[0], O[1], O[2]), R);  // rotation_center = O (Geometry.h:869)
    apply_flags(sp, flags);
    apply_xform(sp, xf, false);
    c->rt->s.addObject(sp);
    int id = (int)c->rt->s.objects.size() - 1;
    if (id == 0) c->rt->s.lumiere = sp;  // Raytracer.cpp:1269
    if (out_id) *out_id = id;
    return PTB_OK;
}

int ptb_add_plane(ptb_ctx* c, const float A[3], const float N[3], const ptb_xform* xf, int flags, int* out_id) {
    if (!c || !A || !N) return PTB_ERR_INVALID;
    Plane* p = new Plane(Vector(A[0], A[1], A[2]), Vector(N[0], N[1], N[2]));
    apply_flags(p, flags);
    apply_xform(p, xf, false);
    c->rt->s.addObject(p);
    if (out_id) *out_id = (int)c->rt->s.objects.size() - 1;
    return PTB_OK;
}

int ptb_add_pointset(ptb_ctx* c, const ptb_pointset* p, const ptb_xform* xf, int flags, int* out_id) {
    if (!c || !p || !p->points || !p->normals || !p->radii || p->n <= 0) return PTB_ERR_INVALID;
    PointSet* ps = new PointSet();                       // what PointSet::init leaves behind, without its file reader / estimate_normals
    const int n = p->n;
    ps->nbcols = 0; ps->is_centered = false;
    Vector center(0., 0., 0.);
    std::map<std::array<float, 7>, std::vector<int>> where;
    for (int i = 0; i < n; i++) {
        ps->vertices.push_back(Vector(p->points[3 * i], p->points[3 * i + 1], p->points[3 * i + 2]));
        ps->normals.push_back(Vector(p->normals[3 * i], p->normals[3 * i + 1], p->normals[3 * i + 2]));
        if (p->colors) ps->colors.push_back(Vector(p->colors[3 * i], p->colors[3 * i + 1], p->colors[3 * i + 2]));
        else ps->colors.push_back(Vector(0.5, 0.5, 0.5));     // build_bvh_recur swaps colors[i] unconditionally; 0.5 grey is the `colors.size() > i` fallback's value
        ps->radius.push_back(p->radii[i]);
        center += ps->vertices[i];
        where[{p->points[3 * i], p->points[3 * i + 1], p->points[3 * i + 2], p->normals[3 * i], p->normals[3 * i + 1], p->normals[3 * i + 2], p->radii[i]}].push_back(i);
    }
    center = center / (float)n;
    ps->rotation_center = center;                        // PointSet.h:113-121
    ps->name = "in-memory";
    ps->build_bvh(0, n);
    std::vector<int>& perm = c->pts_perm[ps];
    perm.resize(n);
    for (int j = 0; j < n; j++) {
        std::vector<int>& v = where[{(float)ps->vertices[j][0], (float)ps->vertices[j][1], (float)ps->vertices[j][2], (float)ps->normals[j][0], (float)ps->normals[j][1],
                                     (float)ps->normals[j][2], (float)ps->radius[j]}];
        perm[j] = v.empty() ? -1 : v.back();
        if (!v.empty()) v.pop_back();
    }
    apply_flags(ps, flags);
    ps->display_edges = (flags & PTB_OBJ_DISPLAY_EDGES) != 0;
    apply_xform(ps, xf, false);
    c->rt->s.addObject(ps);
    if (out_id) *out_id = (int)c->rt->s.objects.size() - 1;
    return PTB_OK;
}

int ptb_add_yarns(ptb_ctx* c, const ptb_yarns* y, const ptb_xform* xf, int flags, int* out_id) {
    if (!c || !y || !y->A || !y->B || !y->R || y->n <= 0) return PTB_ERR_INVALID;
    Yarns* ys = new Yarns();                             // what Yarns(filename) leaves behind (TriangleMesh.h:268-290), without its file reader
    std::map<const Cylinder*, int> handed;
    for (int i = 0; i < y->n; i++) {
        Cylinder* cy = new Cylinder(Vector(y->A[3 * i], y->A[3 * i + 1], y->A[3 * i + 2]), Vector(y->B[3 * i], y->B[3 * i + 1], y->B[3 * i + 2]), y->R[i]);
        ys->cyls.push_back(cy);
        handed[cy] = i;
    }
    ys->build_bvh(&ys->bvh, 0, (int)ys->cyls.size());
    std::vector<int>& perm = c->pts_perm[ys];            // position after build_bvh -> index handed in
    perm.resize(ys->cyls.size());
    for (size_t j = 0; j < ys->cyls.size(); j++) perm[j] = handed[ys->cyls[j]];
    ys->name = "in-memory";
    apply_flags(ys, flags);
    apply_xform(ys, xf, false);
    c->rt->s.addObject(ys);
    if (out_id) *out_id = (int)c->rt->s.objects.size() - 1;
    return PTB_OK;
}

int ptb_add_cylinder(ptb_ctx* c, const float A[3], const float B[3], float R, const ptb_xform* xf, int flags, int* out_id) {
    if (!c || !A || !B) return PTB_ERR_INVALID;
    Cylinder* cy = new Cylinder(Vector(A[0], A[1], A[2]), Vector(B[0], B[1], B[2]), R);
    apply_flags(cy, flags);
    apply_xform(cy, xf, false);
    c->rt->s.addObject(cy);
    if (out_id) *out_id = (int)c->rt->s.objects.size() - 1;
    return PTB_OK;
}

int ptb_add_mesh(ptb_ctx* c, const ptb_mesh* m, const ptb_xform* xf, int flags, int* out_id) {
    if (!c || !m || !m->vertices || !m->tri || m->n_tri <= 0) return PTB_ERR_INVALID;
    auto t0 = std::chrono::steady_clock::now();
    TriMesh* g = new TriMesh();
    g->vertices.resize(m->n_vertices);
    for (int i = 0; i < m->n_vertices; i++) g->vertices[i] = Vector(m->vertices[3 * i], m->vertices[3 * i + 1], m->vertices[3 * i + 2]);
    g->normals.resize(m->n_normals);
    for (int i = 0; i < m->n_normals; i++) g->normals[i] = Vector(m->normals[3 * i], m->normals[3 * i + 1], m->normals[3 * i + 2]);
    g->uvs.resize(m->n_uvs);
    for (int i = 0; i < m->n_uvs; i++) g->uvs[i] = Vector(m->uvs[2 * i], m->uvs[2 * i + 1], 0.f);
    g->indices.resize(m->n_tri);
    for (int i = 0; i < m->n_tri; i++) {
        const int32_t* t = m->tri + 10 * (size_t)i;
        // ctor order: vtx i,j,k, ni,nj,nk, uvi,uvj,uvk, group (TriangleMesh.h:55)
        g->indices[i] = TriangleIndices(t[0], t[1], t[2], t[6], t[7], t[8], t[3], t[4], t[5], t[9]);
    }
    g->bvh_depth = 0;
    g->init(&c->rt->s, "synthetic", m->scaling, Vector(m->offset[0], m->offset[1], m->offset[2]),
            (flags & PTB_OBJ_MIRROR) != 0, NULL, false, false, m->center != 0);
    apply_flags(g, flags);
    g->interp_normals = (flags & PTB_OBJ_FLAT_NORMALS) == 0;
    apply_xform(g, xf, true);
    c->rt->s.addObject(g);
    c->ms_build += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->n_tri += m->n_tri;
    if (out_id) *out_id = (int)c->rt->s.objects.size() - 1;
    return PTB_OK;
}

int ptb_set_group_material(ptb_ctx* c, int obj, int group, const ptb_material* m) {
    if (!c || !m || obj < 0 || obj >= (int)c->rt->s.objects.size() || group < 0) return PTB_ERR_INVALID;
    Object* o = c->rt->s.objects[obj];
    if (m->present & PTB_SLOT_KD) set_slot(o->textures, group, make_tex(m->Kd, 0));
    if (m->present & PTB_SLOT_KS) set_slot(o->specularmap, group, make_tex(m->Ks, 1));
    if (m->present & PTB_SLOT_NE) set_slot(o->roughnessmap, group, make_tex(m->Ne, 4));
    if (m->present & PTB_SLOT_TRANSP) set_slot(o->transparent_map, group, make_tex(m->transp, 5));
    if (m->present & PTB_SLOT_REFR) set_slot(o->refr_index_map, group, make_tex(m->refr, 6));
    if (m->present & PTB_SLOT_NORMAL) set_slot(o->normal_map, group, make_tex(m->normal, 2));
    if (m->present & PTB_SLOT_ALPHA) set_slot(o->alphamap, group, make_tex(m->alpha, 3));
    if (m->present & PTB_SLOT_KSUB) set_slot(o->subsurface, group, make_tex(m->Ksub, 1));
    return PTB_OK;
}

int ptb_add_merl(ptb_ctx* c, const double* table, int* out_merl_id) {
    if (!c || !table) return PTB_ERR_INVALID;
    const size_t n = (size_t)3 * 90 * 90 * 180;
    double* copy = (double*)malloc(n * sizeof(double));
    memcpy(copy, table, n * sizeof(double));
    c->merl.push_back(copy);
    if (out_merl_id) *out_merl_id = (int)c->merl.size() - 1;
    return PTB_OK;
}

int ptb_set_brdf(ptb_ctx* c, int obj, int kind, int merl_id) {
    if (!c || obj < 0 || obj >= (int)c->rt->s.objects.size()) return PTB_ERR_INVALID;
    Object* o = c->rt->s.objects[obj];
    if (kind == PTB_BRDF_PHONG) {
        o->brdf = new PhongBRDF();
    } else if (kind == PTB_BRDF_MERL) {
        if (merl_id < 0 || merl_id >= (int)c->merl.size()) return PTB_ERR_INVALID;
        IsoMERLBRDF* b = new IsoMERLBRDF(std::string(""));  // read_brdf fails on "", leaves data unset
        b->data = c->merl[merl_id];
        o->brdf = b;
    } else {
        return PTB_ERR_UNSUPPORTED;
    }
    return PTB_OK;
}

int ptb_set_envmap(ptb_ctx* c, const uint8_t* rgb, int W, int H) {
    if (!c || c->rt->s.objects.size() < 2) return PTB_ERR_STATE;
    Sphere* dome = dynamic_cast<Sphere*>(c->rt->s.objects[1]);
    if (!dome) return PTB_ERR_STATE;
    if (!rgb || W <= 0 || H <= 0) { dome->has_envmap = false; return PTB_OK; }
    dome->envtex.assign(rgb, rgb + (size_t)W * H * 3);
    dome->envW = W; dome->envH = H;
    dome->has_envmap = true;
    return PTB_OK;
}

int ptb_set_light(ptb_ctx* c, float intensite_lumiere, float envmap_intensity) {
    if (!c) return PTB_ERR_INVALID;
    c->rt->s.intensite_lumiere = intensite_lumiere;
    c->rt->s.envmap_intensity = envmap_intensity;
    return PTB_OK;
}

int ptb_set_keyframes(ptb_ctx* c, int obj, int kind, const float* frames, const float* values, int n) {
    if (!c || obj < 0 || obj >= (int)c->rt->s.objects.size() || n < 0) return PTB_ERR_INVALID;
    Object* o = c->rt->s.objects[obj];
    if (kind == PTB_KEY_SCALE) { o->scale_keyframes.clear(); for (int i = 0; i < n; i++) o->scale_keyframes[frames[i]] = values[i]; }
    else if (kind == PTB_KEY_TRANSLATION) { o->translation_keyframes.clear(); for (int i = 0; i < n; i++) o->translation_keyframes[frames[i]] = Vector(values[3 * i], values[3 * i + 1], values[3 * i + 2]); }
    else if (kind == PTB_KEY_ROTATION) {
        o->rotation_keyframes.clear();
        for (int i = 0; i < n; i++) { Matrix33 m; for (int k = 0; k < 9; k++) m[k] = values[9 * i + k]; o->rotation_keyframes[frames[i]] = m; }
    } else return PTB_ERR_INVALID;
    return PTB_OK;
}

int ptb_set_frame(ptb_ctx* c, float frame) {
    if (!c) return PTB_ERR_INVALID;
    c->frame = (int)frame;            // Scene::current_frame is an int (Geometry.h:1372)
    c->rt->s.current_frame = c->frame;
    return PTB_OK;
}

int ptb_set_fog(ptb_ctx* c, const ptb_fog* f) {
    if (!c || !f) return PTB_ERR_INVALID;
    Scene& s = c->rt->s;
    s.fog_density = f->density; s.fog_absorption = f->absorption;
    s.fog_density_decay = f->density_decay; s.fog_absorption_decay = f->absorption_decay;
    s.fog_type = f->type; s.fog_phase_type = f->phase_type; s.phase_aniso = f->phase_aniso;
    return PTB_OK;
}

int ptb_set_background(ptb_ctx* c, const float* rgb, int W, int H) {
    if (!c) return PTB_ERR_INVALID;
    Scene& s = c->rt->s;
    if (!rgb || W <= 0 || H <= 0) { s.clear_background(); s.backgroundH = 0; return PTB_OK; }
    s.background.assign(rgb, rgb + (size_t)W * H * 3);
    s.backgroundW = W; s.backgroundH = H;
    return PTB_OK;
}

int ptb_commit(ptb_ctx* c) {
    if (!c) return PTB_ERR_INVALID;
    if (c->rt->s.fog_density > 1E-8 && c->rt->s.objects.size() < 3) { c->err = "fog needs object 2 (its translation is the ground level)"; return PTB_ERR_STATE; }
    if (c->rt->s.objects.size() < 2 || !c->rt->s.lumiere) { c->err = "need light (id 0) and dome (id 1)"; return PTB_ERR_STATE; }
    c->committed = true;
    return PTB_OK;
}

static void set_camera(Raytracer* rt, const ptb_camera* cam) {
    rt->cam = Camera(Vector(cam->position[0], cam->position[1], cam->position[2]),
                     Vector(cam->direction[0], cam->direction[1], cam->direction[2]),
                     Vector(cam->up[0], cam->up[1], cam->up[2]));
    rt->cam.fov = cam->fov;
    rt->cam.focus_distance = cam->focus_distance;
    rt->cam.aperture = cam->aperture;
}

static int setup_frame(ptb_ctx* c, const ptb_camera* cam, const ptb_params* p) {
    if (!c || !cam || !p || p->W <= 0 || p->H <= 0 || p->nrays <= 0) return PTB_ERR_INVALID;
    if (!c->committed) return PTB_ERR_STATE;
    Raytracer* rt = c->rt;
    set_camera(rt, cam);
    rt->W = p->W; rt->H = p->H; rt->nrays = p->nrays; rt->nb_bounces = p->nb_bounces;
    rt->sigma_filter = p->sigma_filter; rt->gamma = p->gamma;
    rt->s.double_frustum_start_t = 0; rt->s.current_frame = c->frame;
    ptb_ref_global_seed = p->seed;
    int nt = c->threads > 0 ? c->threads : omp_get_num_procs();
    if (nt > 64) nt = 64;  // engine[64], contribsArray[64] (Vector.h:29, Raytracer.h:114-115)
    omp_set_num_threads(nt);
    return PTB_OK;
}

int ptb_render(ptb_ctx* c, const ptb_camera* cam, const ptb_params* p,
               float* imagedouble, float* sample_count, uint8_t* image, ptb_stats* stats) {
    int rc = setup_frame(c, cam, p);
    if (rc) return rc;
    if (p->shard_count > 1) return PTB_ERR_UNSUPPORTED;
    Raytracer* rt = c->rt;
    memset(ptb_ref_cnt, 0, sizeof(ptb_ref_cnt));
    auto t0 = std::chrono::steady_clock::now();
    rt->render_image_nopreviz();
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    size_t n = (size_t)p->W * p->H;
    if (imagedouble) memcpy(imagedouble, &rt->imagedouble[0], n * 3 * sizeof(float));
    if (sample_count) memcpy(sample_count, &rt->sample_count[0], n * sizeof(float));
    if (image) memcpy(image, &rt->image[0], n * 3);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->samples = (uint64_t)n * p->nrays;
        for (int t = 0; t < 64; t++) { stats->rays_closest += ptb_ref_cnt[t][0]; stats->rays_shadow += ptb_ref_cnt[t][1]; }
        stats->ms_wall = ms;
        stats->ms_device = 0;
    }
    return PTB_OK;
}

// render_image_nopreviz with has_denoiser = true: without OIDN compiled in, the reference stops after the normalisation loop
// (Raytracer.cpp:1687-1694; build_ref.py patch 3 closes the brace the OIDN block would have closed)
int ptb_render_denoiser_inputs(ptb_ctx* c, const ptb_camera* cam, const ptb_params* p, float* imagedouble, float* sample_count, float* albedoImage,
                               float* normalImage, float* first_hit_normal, ptb_stats* stats) {
    int rc = setup_frame(c, cam, p);
    if (rc) return rc;
    Raytracer* rt = c->rt;
    memset(ptb_ref_cnt, 0, sizeof(ptb_ref_cnt));
    rt->has_denoiser = true;
    rt->render_image_nopreviz();
    rt->has_denoiser = false;
    size_t n = (size_t)p->W * p->H;
    if (imagedouble) memcpy(imagedouble, &rt->imagedouble[0], n * 3 * sizeof(float));
    if (sample_count) memcpy(sample_count, &rt->sample_count[0], n * sizeof(float));
    if (albedoImage) memcpy(albedoImage, &rt->albedoImage[0], n * 3 * sizeof(float));
    if (normalImage) memcpy(normalImage, &rt->normalImage[0], n * 3 * sizeof(float));
    if (first_hit_normal) for (size_t i = 0; i < n * 3; i++) first_hit_normal[i] = NAN;   // the reference never forms this quantity (1680-1682 sum colours)
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->samples = (uint64_t)n * p->nrays;
        for (int t = 0; t < 64; t++) { stats->rays_closest += ptb_ref_cnt[t][0]; stats->rays_shadow += ptb_ref_cnt[t][1]; }
    }
    return PTB_OK;
}

// Raytracer::render_image cannot be paused from outside; with per-(pixel,sample) streams (build_ref.py patch 5c) the first k passes
// of a longer render ARE a k-pass render, so a session re-renders `passes so far` on read.
static ptb_camera g_prog_cam; static ptb_params g_prog_p; static int g_prog_iter = -1;
int ptb_progressive_begin(ptb_ctx* c, const ptb_camera* cam, const ptb_params* p) {
    int rc = setup_frame(c, cam, p);
    if (rc) return rc;
    g_prog_cam = *cam; g_prog_p = *p; g_prog_iter = 0;
    return PTB_OK;
}
int ptb_progressive_pass(ptb_ctx* c, int n_spp, ptb_stats* stats) {
    if (!c || g_prog_iter < 0) return PTB_ERR_STATE;
    int n = n_spp < g_prog_p.nrays - g_prog_iter ? n_spp : g_prog_p.nrays - g_prog_iter;
    if (n > 0) g_prog_iter += n;
    if (stats) memset(stats, 0, sizeof(*stats));
    return PTB_OK;
}
int ptb_progressive_read(ptb_ctx* c, float* imagedouble, float* sample_count, uint8_t* image, float* imagedouble_lowres, int32_t* current_nb_rays) {
    if (!c || g_prog_iter < 0) return PTB_ERR_STATE;
    ptb_params p = g_prog_p;
    p.nrays = g_prog_iter > 0 ? g_prog_iter : 1;
    int rc = setup_frame(c, &g_prog_cam, &p);
    if (rc) return rc;
    Raytracer* rt = c->rt;
    if (g_prog_iter == 0) rt->nrays = 0;
    rt->stopped = false;
    rt->render_image();
    size_t n = (size_t)p.W * p.H;
    if (imagedouble) memcpy(imagedouble, &rt->imagedouble[0], n * 3 * sizeof(float));
    if (sample_count) memcpy(sample_count, &rt->sample_count[0], n * sizeof(float));
    if (image) memcpy(image, &rt->image[0], n * 3);
    if (imagedouble_lowres) memcpy(imagedouble_lowres, &rt->imagedouble_lowres[0], (size_t)rt->Wlr * rt->Hlr * 3 * sizeof(float));
    if (current_nb_rays) *current_nb_rays = g_prog_iter;
    return PTB_OK;
}

int ptb_render_accum(ptb_ctx*, const ptb_camera*, const ptb_params*, float*, ptb_stats*) { return PTB_ERR_UNSUPPORTED; }
int ptb_resolve(ptb_ctx*, const float*, int, int, float, float*, float*, uint8_t*) { return PTB_ERR_UNSUPPORTED; }
int ptb_shard_pack_size(const ptb_params*, int, int64_t*) { return PTB_ERR_UNSUPPORTED; }
int ptb_shard_pack(ptb_ctx*, const ptb_params*, int, const float*, float*) { return PTB_ERR_UNSUPPORTED; }
int ptb_shard_unpack_add(ptb_ctx*, const ptb_params*, int, const float*, float*) { return PTB_ERR_UNSUPPORTED; }

int ptb_primary_ids(ptb_ctx* c, const ptb_camera* cam, int W, int H, int32_t* obj_id, int32_t* tri_id, float* tout) {
    if (!c || !cam || W <= 0 || H <= 0) return PTB_ERR_INVALID;
    if (!c->committed) return PTB_ERR_STATE;
    Raytracer* rt = c->rt;
    set_camera(rt, cam);
    rt->s.current_frame = c->frame;
    rt->s.prepare_render(false);
    // the picking query, mainApp.h:686-692
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < H; i++) {
        for (int j = 0; j < W; j++) {
            Ray r = rt->cam.generateDirection(0, i, j, 0, 0, 0, 0, 0, W, H);
            Vector P; MaterialValues mat; int id = -1, tri = -1; float t;
            bool hit = rt->s.intersection(r, P, id, t, mat, tri);
            int32_t oid = -1, tid = -1;
            if (hit) {
                oid = id;
                TriMesh* g = rt->s.castToMesh[id];
                if (g && tri >= 0) tid = g->permuted_triangle_index[tri];
                auto it = c->pts_perm.find(rt->s.objects[id]);
                if (it != c->pts_perm.end() && tri >= 0 && tri < (int)it->second.size()) tid = it->second[tri];
            }
            if (obj_id) obj_id[(size_t)i * W + j] = oid;
            if (tri_id) tri_id[(size_t)i * W + j] = tid;
            if (tout) tout[(size_t)i * W + j] = hit ? t : -1.f;
        }
    }
    return PTB_OK;
}

int ptb_set_option(ptb_ctx* c, int option, int64_t value) {
    if (!c) return PTB_ERR_INVALID;
    if (option == ORC_OPT_THREADS) { c->threads = (int)value; return PTB_OK; }
    return PTB_OK;  // GPU-only options are ignored
}

int ptb_get_scene_info(const ptb_ctx* c, ptb_scene_info* info) {
    if (!c || !info) return PTB_ERR_INVALID;
    memset(info, 0, sizeof(*info));
    info->n_triangles = c->n_tri;
    info->n_objects = (int)c->rt->s.objects.size();
    info->ms_bvh_build = c->ms_build;
    for (size_t i = 0; i < c->rt->s.objects.size(); i++) {
        TriMesh* g = c->rt->s.castToMesh[i];
        if (g) { info->n_bvh_nodes += g->bvh.nodes.size(); if (g->bvh_depth > info->bvh_depth) info->bvh_depth = g->bvh_depth; }
    }
    info->bytes_nodes = info->n_bvh_nodes * (int64_t)sizeof(BVHNodes);
    info->bytes_triangles = info->n_triangles * (int64_t)sizeof(Triangle);
    return PTB_OK;
}

int ptb_get_kernel_times(const ptb_ctx*, ptb_kernel_times*) { return PTB_ERR_UNSUPPORTED; }
int ptb_kat(ptb_ctx* c, int which, const ptb_camera* cam, int W, int H,
            const double* in, int n, int is, double* out, int os) {
    if (!c || !in || !out) return PTB_ERR_INVALID;
    Raytracer* rt = c->rt;
    if (cam) set_camera(rt, cam);
    if (which == PTB_KAT_RANDOM_PER_PIXEL || which == PTB_KAT_FILTER_RATIO) {
        if (!c->committed) return PTB_ERR_STATE;
        rt->W = W; rt->H = H;
        if (which == PTB_KAT_FILTER_RATIO) { rt->sigma_filter = (float)in[2]; rt->lastfilter = -1; }
        rt->prepare_render(0);
    }
    for (int k = 0; k < n; k++) {
        const double* a = in + (size_t)k * is;
        double* o = out + (size_t)k * os;
        switch (which) {
        case PTB_KAT_PCG32: {
            pcg32 e((uint64_t)a[0], (uint64_t)a[1]);
            for (int q = 0; q < 4; q++) o[q] = (double)e();
        } break;
        case PTB_KAT_LATTICE: {
            Vector v = extensibleLattice2d((uint32_t)a[0]);
            o[0] = v[0]; o[1] = v[1];
        } break;
        case PTB_KAT_CAMERA: {
            Ray r = rt->cam.generateDirection(0, (int)a[0], (int)a[1], 0, (float)a[2], (float)a[3], (float)a[4], (float)a[5], W, H);
            for (int q = 0; q < 3; q++) { o[q] = r.origin[q]; o[3 + q] = r.direction[q]; }
        } break;
        case PTB_KAT_RANDOM_COS: {
            Vector v = random_cos(Vector((float)a[0], (float)a[1], (float)a[2]), (float)a[3], (float)a[4]);
            for (int q = 0; q < 3; q++) o[q] = v[q];
        } break;
        case PTB_KAT_RANDOM_PHONG: {
            Vector v = PhongBRDF::random_Phong(Vector((float)a[0], (float)a[1], (float)a[2]), (float)a[3], (float)a[4], (float)a[5]);
            for (int q = 0; q < 3; q++) o[q] = v[q];
        } break;
        case PTB_KAT_PHONG_EVAL: {
            MaterialValues m;
            m.Kd = Vector((float)a[0], (float)a[1], (float)a[2]);
            m.Ks = Vector((float)a[3], (float)a[4], (float)a[5]);
            m.Ne = Vector((float)a[6], (float)a[7], (float)a[8]);
            PhongBRDF b;
            Vector v = b.eval(m, Vector((float)a[9], (float)a[10], (float)a[11]), Vector((float)a[12], (float)a[13], (float)a[14]),
                              Vector((float)a[15], (float)a[16], (float)a[17]));
            for (int q = 0; q < 3; q++) o[q] = v[q];
        } break;
        case PTB_KAT_MERL_EVAL: {
            if (c->merl.empty()) return PTB_ERR_STATE;
            IsoMERLBRDF b(std::string(""));
            b.data = c->merl[0];
            MaterialValues m;
            Vector v = b.eval(m, Vector((float)a[0], (float)a[1], (float)a[2]), Vector((float)a[3], (float)a[4], (float)a[5]),
                              Vector((float)a[6], (float)a[7], (float)a[8]));
            for (int q = 0; q < 3; q++) o[q] = v[q];
        } break;
        case PTB_KAT_FAST_EXP: o[0] = fast_exp(a[0]); break;
        case PTB_KAT_FAST_NORMALIZE: {
            Vector v((float)a[0], (float)a[1], (float)a[2]);
            v.fast_normalize();
            for (int q = 0; q < 3; q++) o[q] = v[q];
        } break;
        case PTB_KAT_RANDOM_PER_PIXEL: {
            size_t p = (size_t)a[0];
            o[0] = rt->randomPerPixel[p][0]; o[1] = rt->randomPerPixel[p][1];
        } break;
        case PTB_KAT_FILTER_RATIO: {
            int i = (int)a[0], j = (int)a[1];
            int fs = rt->filter_size, ftw = rt->filter_total_width;
            int bmin_i = std::max(0, i - fs), bmax_i = std::min(i + fs, H - 1);
            int bmin_j = std::max(0, j - fs), bmax_j = std::min(j + fs, W - 1);
            float ratio = 1.f / sum_area_table(&rt->filter_integral[0], ftw, bmin_i - i + fs, bmax_i - i + fs, bmin_j - j + fs, bmax_j - j + fs);
            o[0] = ratio;
        } break;
        default: return PTB_ERR_UNSUPPORTED;
        }
    }
    return PTB_OK;
}


// ---- the reference's own FILE READERS behind the accessor shapes of include/ptb_sceneio.h (prefix ref_) ----------------
// Used by tests/golden/make_golden.py and tests/test_scene_io.py to pin pathtracer_b200/csrc/scene_io.cpp: the same Python
// comparison code walks the product's handles and these.
#include "ptb_sceneio.h"

struct ref_meshfile_t { TriMesh* g; bool has_materials; std::vector<float> v, n, uv, vc; std::vector<int32_t> tri; };

static void ref_copy_str(char* dst, const std::string& s) {
    size_t n = s.size() < PTB_PATH_MAX - 1 ? s.size() : PTB_PATH_MAX - 1;
    memcpy(dst, s.data(), n); dst[n] = 0;
}
static void ref_slot_from_tex(const Texture& t, ptb_slot* out) {
    // a slot whose image did not load keeps W == 0: report it as constant ("" file) like the product's reader resolves it
    ref_copy_str(out->file, (t.W > 0) ? t.filename : std::string());
    for (int k = 0; k < 3; k++) out->mult[k] = t.multiplier[k];
}
static const std::vector<Texture>* ref_slot_vector(const Object* o, int kind) {
    switch (kind) {
    case PTB_KIND_KD: return &o->textures;
    case PTB_KIND_NORMAL: return &o->normal_map;
    case PTB_KIND_SUBSURF: return &o->subsurface;
    case PTB_KIND_KS: return &o->specularmap;
    case PTB_KIND_ALPHA: return &o->alphamap;
    case PTB_KIND_NE: return &o->roughnessmap;
    case PTB_KIND_TRANSP: return &o->transparent_map;
    case PTB_KIND_REFR: return &o->refr_index_map;
    }
    return NULL;
}

int ref_image_load(const char* path, uint8_t** rgb, int32_t* W, int32_t* H) {
    std::vector<unsigned char> v; size_t w = 0, h = 0;
    if (!load_image(path, v, w, h)) return PTB_ERR_INVALID;
    *rgb = (uint8_t*)malloc(v.size()); memcpy(*rgb, v.data(), v.size()); *W = (int32_t)w; *H = (int32_t)h;
    return PTB_OK;
}
void ref_image_free(void* p) { free(p); }
int ref_texture_load(const char* path, int kind, float** values, int32_t* W, int32_t* H) {
    Texture t;
    t.W = 0; t.H = 0;
    if (kind == 0) t.loadColors(path); else t.loadNormals(path);
    if (t.W == 0) return PTB_ERR_INVALID;
    *values = (float*)malloc(t.values.size() * sizeof(float)); memcpy(*values, t.values.data(), t.values.size() * sizeof(float));
    *W = (int32_t)t.W; *H = (int32_t)t.H;
    return PTB_OK;
}

// `new Yarns(file)`: its constructor reads the file and builds its BVH, which reorders `cyls`; the segments come back in THAT order
// (tests compare them as a set).  The constructor opens with "r+" and does not check the result: probe first.
int ref_yarnfile_read(const char* path, float** A, float** B, float** R, int32_t* n_segments) {
    FILE* probe = fopen(path, "r+");
    if (!probe) return PTB_ERR_INVALID;
    fclose(probe);
    Yarns* y = new Yarns(path);
    const size_t n = y->cyls.size();
    *A = (float*)malloc(3 * n * sizeof(float)); *B = (float*)malloc(3 * n * sizeof(float)); *R = (float*)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) { (*A)[3 * i + k] = y->cyls[i]->A[k]; (*B)[3 * i + k] = y->cyls[i]->B[k]; }
        (*R)[i] = y->cyls[i]->R;
    }
    *n_segments = (int32_t)n;
    return PTB_OK;
}
void ref_yarnfile_free(void* p) { free(p); }

int ref_meshfile_read(const char* path, int load_textures, void** out) {
    ref_meshfile_t* m = new ref_meshfile_t();
    m->g = new TriMesh();
    m->has_materials = load_textures != 0;
    std::string lo(path); for (size_t i = 0; i < lo.size(); i++) lo[i] = tolower(lo[i]);
    FILE* probe = fopen(path, "r");
    if (!probe) { delete m; return PTB_ERR_INVALID; }
    fclose(probe);
    if (lo.find(".off") != std::string::npos) m->g->readOFF(path);        // TriMesh::init's dispatch (TriangleMesh.cpp:729-741)
    else if (lo.find(".obj") != std::string::npos) m->g->readOBJ(path, load_textures != 0);
    else { delete m; return PTB_ERR_UNSUPPORTED; }
    TriMesh* g = m->g;
    for (size_t i = 0; i < g->vertices.size(); i++) for (int k = 0; k < 3; k++) m->v.push_back(g->vertices[i][k]);
    for (size_t i = 0; i < g->normals.size(); i++) for (int k = 0; k < 3; k++) m->n.push_back(g->normals[i][k]);
    for (size_t i = 0; i < g->uvs.size(); i++) for (int k = 0; k < 2; k++) m->uv.push_back(g->uvs[i][k]);
    for (size_t i = 0; i < g->vertexcolors.size(); i++) for (int k = 0; k < 3; k++) m->vc.push_back(g->vertexcolors[i][k]);
    for (size_t i = 0; i < g->indices.size(); i++) {
        const TriangleIndices& t = g->indices[i];
        int32_t r[10] = {t.vtxi, t.vtxj, t.vtxk, t.uvi, t.uvj, t.uvk, t.ni, t.nj, t.nk, t.group};
        m->tri.insert(m->tri.end(), r, r + 10);
    }
    *out = m;
    return PTB_OK;
}
void ref_meshfile_free(void* h) { delete (ref_meshfile_t*)h; }
int ref_meshfile_get(const void* h, ptb_meshfile_info* o) {
    const ref_meshfile_t* m = (const ref_meshfile_t*)h;
    o->vertices = m->v.data(); o->n_vertices = (int32_t)(m->v.size() / 3);
    o->normals = m->n.data(); o->n_normals = (int32_t)(m->n.size() / 3);
    o->uvs = m->uv.data(); o->n_uvs = (int32_t)(m->uv.size() / 2);
    o->vertex_colors = m->vc.data(); o->n_vertex_colors = (int32_t)(m->vc.size() / 3);
    o->tri = m->tri.data(); o->n_tri = (int32_t)(m->tri.size() / 10);
    o->n_groups = (int32_t)m->g->groupNames.size();
    o->has_materials = m->has_materials ? 1 : 0;
    return PTB_OK;
}
int ref_meshfile_group_name(const void* h, int group, char name[PTB_PATH_MAX]) {
    const ref_meshfile_t* m = (const ref_meshfile_t*)h;
    for (std::map<std::string, int>::const_iterator it = m->g->groupNames.begin(); it != m->g->groupNames.end(); ++it)
        if (it->second == group) { ref_copy_str(name, it->first); return PTB_OK; }
    return PTB_ERR_INVALID;
}
int ref_meshfile_group_slot(const void* h, int group, int kind, ptb_slot* out) {
    const ref_meshfile_t* m = (const ref_meshfile_t*)h;
    const std::vector<Texture>* v = ref_slot_vector(m->g, kind);
    if (!v || group < 0 || group >= (int)v->size()) return PTB_ERR_INVALID;
    ref_slot_from_tex((*v)[group], out);
    return PTB_OK;
}

int ref_scn_load(const char* path, const char* replaced, void** out) {
    FILE* probe = fopen(path, "r");
    if (!probe) return PTB_ERR_INVALID;
    fclose(probe);
    ptb_ctx* c = NULL;
    ptb_create(0, &c);
    c->rt->load_scene(path, replaced);
    *out = c;
    return PTB_OK;
}
void ref_scn_free(void* h) { ptb_destroy((ptb_ctx*)h); }
int ref_scn_get_header(const void* hh, ptb_scn_header* h) {
    const Raytracer* rt = ((const ptb_ctx*)hh)->rt;
    memset(h, 0, sizeof(*h));
    h->W = rt->W; h->H = rt->H; h->nrays = rt->nrays; h->nbframes = rt->s.nbframes; h->nb_bounces = rt->nb_bounces;
    h->has_denoiser = rt->has_denoiser; h->is_lenticular = rt->cam.is_lenticular; h->n_objects = (int32_t)rt->s.objects.size();
    for (int k = 0; k < 3; k++) { h->cam.position[k] = rt->cam.position[k]; h->cam.direction[k] = rt->cam.direction[k]; h->cam.up[k] = rt->cam.up[k]; }
    h->cam.fov = rt->cam.fov; h->cam.focus_distance = rt->cam.focus_distance; h->cam.aperture = rt->cam.aperture;
    h->sigma_filter = rt->sigma_filter; h->gamma = rt->gamma; h->intensite_lumiere = rt->s.intensite_lumiere; h->envmap_intensity = rt->s.envmap_intensity;
    h->fog_density = rt->s.fog_density; h->fog_absorption = rt->s.fog_absorption; h->fog_density_decay = rt->s.fog_density_decay;
    h->fog_absorption_decay = rt->s.fog_absorption_decay; h->fog_type = rt->s.fog_type; h->fog_phase_type = rt->s.fog_phase_type;
    h->double_frustum_start_t = rt->s.double_frustum_start_t;
    if (rt->s.backgroundW > 0) ref_copy_str(h->background, rt->s.backgroundfilename);
    return PTB_OK;
}
int ref_scn_get_object(const void* hh, int i, ptb_scn_object* o) {
    const Raytracer* rt = ((const ptb_ctx*)hh)->rt;
    if (i < 0 || i >= (int)rt->s.objects.size()) return PTB_ERR_INVALID;
    Object* b = rt->s.objects[i];
    memset(o, 0, sizeof(*o));
    o->type = b->type == OT_TRIMESH ? PTB_SCN_MESH : b->type == OT_SPHERE ? PTB_SCN_SPHERE : b->type == OT_PLANE ? PTB_SCN_PLANE : PTB_SCN_POINTSET;
    ref_copy_str(o->name, b->name);
    o->miroir = b->miroir; o->ghost = b->ghost; o->display_edges = b->display_edges; o->interp_normals = b->interp_normals; o->flip_normals = b->flip_normals;
    o->n_keyframes = (int32_t)b->translation_keyframes.size();
    o->xform.scale = b->get_scale(0, false);
    Matrix33 R = b->get_rotation(0, false);
    Vector T = b->get_translation(0, false);
    for (int k = 0; k < 9; k++) o->xform.rotation[k] = R[k];
    for (int k = 0; k < 3; k++) { o->xform.rotation_center[k] = b->rotation_center[k]; o->xform.translation[k] = T[k]; }
    for (int k = 0; k < PTB_N_KINDS; k++) o->n_slots[k] = (int32_t)ref_slot_vector(b, k)->size();
    if (Sphere* sp = dynamic_cast<Sphere*>(b)) {
        o->is_envmap = sp->has_envmap; if (sp->has_envmap) ref_copy_str(o->envmap, sp->envmapfilename);
        for (int k = 0; k < 3; k++) o->O[k] = sp->O[k];
        o->R = sp->R;
    } else if (Plane* pl = dynamic_cast<Plane*>(b)) {
        for (int k = 0; k < 3; k++) { o->A[k] = pl->A[k]; o->N[k] = pl->vecN[k]; }
    } else if (TriMesh* g = dynamic_cast<TriMesh*>(b)) {
        o->is_centered = g->is_centered; o->has_csv = g->csv_file.size() != 0; ref_copy_str(o->csv_file, g->csv_file);
    }
    return PTB_OK;
}
int ref_scn_get_keyframes(const void* hh, int i, int kind, float* frames, float* values, int cap) {
    const Raytracer* rt = ((const ptb_ctx*)hh)->rt;
    if (i < 0 || i >= (int)rt->s.objects.size()) return PTB_ERR_INVALID;
    const Object* b = rt->s.objects[i];
    int n = 0;
    if (kind == PTB_KEY_SCALE) for (std::map<float, float>::const_iterator it = b->scale_keyframes.begin(); it != b->scale_keyframes.end(); ++it, ++n) { if (n < cap) { if (frames) frames[n] = it->first; if (values) values[n] = it->second; } }
    else if (kind == PTB_KEY_TRANSLATION) for (std::map<float, Vector>::const_iterator it = b->translation_keyframes.begin(); it != b->translation_keyframes.end(); ++it, ++n) { if (n < cap) { if (frames) frames[n] = it->first; if (values) for (int k = 0; k < 3; k++) values[3 * n + k] = it->second[k]; } }
    else if (kind == PTB_KEY_ROTATION) for (std::map<float, Matrix33>::const_iterator it = b->rotation_keyframes.begin(); it != b->rotation_keyframes.end(); ++it, ++n) { if (n < cap) { if (frames) frames[n] = it->first; if (values) for (int k = 0; k < 9; k++) values[9 * n + k] = it->second[k]; } }
    else return PTB_ERR_INVALID;
    return n;
}
int ref_scn_get_slot(const void* hh, int i, int kind, int idx, ptb_slot* out) {
    const Raytracer* rt = ((const ptb_ctx*)hh)->rt;
    if (i < 0 || i >= (int)rt->s.objects.size()) return PTB_ERR_INVALID;
    const std::vector<Texture>* v = ref_slot_vector(rt->s.objects[i], kind);
    if (!v || idx < 0 || idx >= (int)v->size()) return PTB_ERR_INVALID;
    ref_slot_from_tex((*v)[idx], out);
    return PTB_OK;
}
// Raytracer::load_scene + hand the loaded Raytracer to the render entry points of this driver
int ref_load_scene(ptb_ctx* c, const char* path, const char* replaced, ptb_camera* cam, ptb_params* p) {
    if (!c || !path) return PTB_ERR_INVALID;
    FILE* probe = fopen(path, "r");
    if (!probe) return PTB_ERR_INVALID;
    fclose(probe);
    Raytracer* rt = c->rt;
    rt->load_scene(path, replaced);
    ptb_scn_header h;
    ref_scn_get_header(c, &h);
    if (cam) *cam = h.cam;
    if (p) { memset(p, 0, sizeof(*p)); p->W = h.W; p->H = h.H; p->nrays = h.nrays; p->nb_bounces = h.nb_bounces; p->sigma_filter = h.sigma_filter; p->gamma = h.gamma; p->shard_count = 1; }
    for (size_t i = 0; i < rt->s.objects.size(); i++) if (TriMesh* g = rt->s.castToMesh[i]) c->n_tri += (long long)g->indices.size();
    return PTB_OK;
}
const char* ref_sceneio_last_error(void) { return "reference reader failed (the reference reports no reason)"; }
int ref_scn_save(const void* hh, const char* path) { ((const ptb_ctx*)hh)->rt->save_scene(path); return PTB_OK; }

}  // extern "C"
