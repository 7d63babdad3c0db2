/*
 * ptb_sceneio.h — C-ABI of the scene-file readers that sit in front of ptb200.h.
 *
 * These replace the reference's file-side callers of the hot path: `Raytracer::load_scene`
 * (Raytracer.cpp:1149-1236), `Object::load_from_file` and the per-type `create_from_file`
 * (Geometry.h:518-662, 887-911, 1203-1213; TriangleMesh.h:143-167), `TriMesh::readOBJ` with its MTL
 * reader (TriangleMesh.cpp:240-569), `TriMesh::readOFF` (107-130) and `load_image` + `Texture::loadColors /
 * loadNormals` (utils.cpp:98-170, BRDF.h:393-419).  Host code only: nothing here touches the GPU, so the
 * functions work on a machine without one.  They live in libptb200.so next to the renderer.
 *
 * Everything is returned as the arrays / fields the reference's readers leave in memory, so the same
 * parsed scene can be handed to ptb200.h or to a CPU checker.  Errors: negative PTB_ERR_* code, text
 * through ptb_sceneio_last_error() (thread-local).
 */
#ifndef PTB_SCENEIO_H
#define PTB_SCENEIO_H

#include "ptb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_PATH_MAX 512

const char* ptb_sceneio_last_error(void);

/* ---- images --------------------------------------------------------------------------------------------- */
/* replaces: load_image<unsigned char> (utils.cpp:98-170), stb path: 8-bit RGB, 3 channels forced, rows flipped
 * (first row of the result = last row of the file).  Formats: PNG (8/16-bit, all colour types, non-interlaced),
 * BMP (8-bit palette, 24, 32 bit), TGA (types 2,3,10,11), PNM (P5/P6, maxval <= 255).  JPEG is not decoded
 * (PTB_ERR_UNSUPPORTED).  The buffer is released with ptb_image_free. */
int  ptb_image_load(const char* path, uint8_t** rgb, int32_t* W, int32_t* H);
void ptb_image_free(void* p);
/* replaces: Texture::loadColors (kind 0: v/255 then powf(.,2.2f), BRDF.h:393-404) and Texture::loadNormals
 * (kind 1: (v-128) normalised per texel, BRDF.h:406-419): the float W*H*3 `Texture::values` array. */
int  ptb_texture_load(const char* path, int kind, float** values, int32_t* W, int32_t* H);

/* ---- mesh files --------------------------------------------------------------------------------------- */
/* Slot kinds, in the order Object::save_to_file writes them (Geometry.h:476-515). */
#define PTB_KIND_KD        0   /* textures          */
#define PTB_KIND_NORMAL    1   /* normal_map        */
#define PTB_KIND_SUBSURF   2   /* subsurface        */
#define PTB_KIND_KS        3   /* specularmap       */
#define PTB_KIND_ALPHA     4   /* alphamap          */
#define PTB_KIND_NE        5   /* roughnessmap      */
#define PTB_KIND_TRANSP    6   /* transparent_map   */
#define PTB_KIND_REFR      7   /* refr_index_map    */
#define PTB_N_KINDS        8

typedef struct ptb_slot {             /* one `Texture` of an Object slot vector as a reader leaves it */
    char    file[PTB_PATH_MAX];       /* image to load ("" = none: the slot is the constant `mult`) */
    float   mult[3];                  /* Texture::multiplier */
} ptb_slot;

typedef struct ptb_meshfile ptb_meshfile;
typedef struct ptb_meshfile_info {
    const float*   vertices;      int32_t n_vertices;      /* n x 3, file axes */
    const float*   normals;       int32_t n_normals;       /* n x 3 */
    const float*   uvs;           int32_t n_uvs;           /* n x 2 */
    const float*   vertex_colors; int32_t n_vertex_colors; /* n x 3 (OBJ "v x y z r g b"), clamped to [0,1] */
    const int32_t* tri;           int32_t n_tri;           /* n x 10 {vtx i,j,k, uv i,j,k, normal i,j,k, group}, -1 absent */
    int32_t        n_groups;                               /* TriMesh::groupNames.size() */
    int32_t        has_materials;                          /* 1: read with load_textures, ptb_meshfile_group_slot is valid */
} ptb_meshfile_info;
/* replaces: TriMesh::readOBJ(obj, load_textures) / readOFF, chosen by extension like TriMesh::init
 * (TriangleMesh.cpp:729-741).  Faces are fan-triangulated (390-458), negative indices resolved against
 * the counts read so far, groups numbered by first `usemtl` appearance; with load_textures the per-group
 * defaults of 481-490 are created and the MTL's Kd/Ks/Ns/map_* override them (492-565). */
int  ptb_meshfile_read(const char* path, int load_textures, ptb_meshfile** out);
void ptb_meshfile_free(ptb_meshfile*);
int  ptb_meshfile_get(const ptb_meshfile*, ptb_meshfile_info* out);
int  ptb_meshfile_group_name(const ptb_meshfile*, int group, char name[PTB_PATH_MAX]);
int  ptb_meshfile_group_slot(const ptb_meshfile*, int group, int kind, ptb_slot* out);

/* ---- .yarn files ---------------------------------------------------------------------------------------- */
/* replaces: the reader inside `Yarns::Yarns(const char* filename)` (TriangleMesh.h:268-288): `nbyarns`, then per yarn `nbsegments`
 * followed by that many points "x y z"; consecutive points of a yarn, times 50, become the end points of one segment of radius 0.1.
 * Returns malloc'ed n x 3 / n x 3 / n arrays ready for ptb_add_yarns (free with ptb_yarnfile_free), in file order.  A count that
 * the file does not honour is an error here (the reference's fscanf loop would push whatever the last conversion left). */
int  ptb_yarnfile_read(const char* path, float** A, float** B, float** R, int32_t* n_segments);
void ptb_yarnfile_free(void* p);

/* ---- .scn files ----------------------------------------------------------------------------------------- */
#define PTB_SCN_MESH      0
#define PTB_SCN_SPHERE    1
#define PTB_SCN_PLANE     2
#define PTB_SCN_POINTSET  3           /* parsed far enough to skip; not renderable here */

typedef struct ptb_scn_header {       /* the Raytracer / Scene / Camera fields load_scene fills */
    int32_t    W, H, nrays, nbframes, nb_bounces, has_denoiser, is_lenticular, n_objects;
    ptb_camera cam;
    float      sigma_filter, gamma, intensite_lumiere, envmap_intensity;
    float      fog_density, fog_absorption, fog_density_decay, fog_absorption_decay;
    int32_t    fog_type, fog_phase_type;
    float      double_frustum_start_t;
    char       background[PTB_PATH_MAX];   /* "" = none */
} ptb_scn_header;

typedef struct ptb_scn_object {
    int32_t   type;                   /* PTB_SCN_* */
    char      name[PTB_PATH_MAX];     /* Object::name: the mesh file for meshes */
    int32_t   miroir, ghost, display_edges, interp_normals, flip_normals;
    int32_t   n_keyframes;            /* nb_transforms; the placement below is evaluated at frame 0 (Slerp between rotation keys) */
    ptb_xform xform;                  /* scale, mat_rotation, rotation_center, max_translation */
    int32_t   n_slots[PTB_N_KINDS];   /* length of each slot vector */
    int32_t   is_envmap;  char envmap[PTB_PATH_MAX];  float O[3], R;   /* sphere */
    float     A[3], N[3];                                              /* plane  */
    int32_t   is_centered, has_csv;  char csv_file[PTB_PATH_MAX];      /* mesh   */
} ptb_scn_object;

typedef struct ptb_scn ptb_scn;
/* replaces: the parsing half of Raytracer::load_scene.  `replaced_names` substitutes the first '#' of
 * every object name (Geometry.h:524-526); NULL leaves names alone. */
int  ptb_scn_load(const char* path, const char* replaced_names, ptb_scn** out);
void ptb_scn_free(ptb_scn*);
int  ptb_scn_get_header(const ptb_scn*, ptb_scn_header* out);
int  ptb_scn_get_object(const ptb_scn*, int obj, ptb_scn_object* out);
int  ptb_scn_get_slot(const ptb_scn*, int obj, int kind, int idx, ptb_slot* out);
/* The object's keyframe maps (Object::{scale,translation,rotation}_keyframes, Geometry.h:318-320, nb_transforms rows each):
 * kind = PTB_KEY_SCALE / _TRANSLATION / _ROTATION of ptb200.h, values n x 1 / 3 / 9, frames ascending.  Returns the number of keys
 * (writes at most `cap`); frames / values may be NULL to query the count. */
int  ptb_scn_get_keyframes(const ptb_scn*, int obj, int kind, float* frames, float* values, int cap);
/* replaces: Raytracer::save_scene (Raytracer.cpp:1096-1146) for a parsed scene (round-trip tests, tools). */
int  ptb_scn_save(const ptb_scn*, const char* path);

/* replaces: Raytracer::load_scene end to end for a C/C++ caller: parses `path`, reads the meshes, textures
 * and the dome's environment map it names (paths as written, else relative to the .scn's directory), feeds
 * them to `ctx` through ptb_add_* / ptb_set_* and fills the camera and frame parameters.  The caller then
 * calls ptb_commit + ptb_render.  Features the renderer does not implement (PointSet objects, VRML meshes,
 * vertex colours, environment maps on spheres other than object 1) fail with PTB_ERR_UNSUPPORTED
 * instead of rendering something else. */
int  ptb_load_scene(ptb_ctx* ctx, const char* path, const char* replaced_names, ptb_camera* cam, ptb_params* params);

#ifdef __cplusplus
}
#endif
#endif /* PTB_SCENEIO_H */
