/* oracle/prefix.h — TEST INFRASTRUCTURE.  Renames the ptb_* entry points of include/ptb200.h so the
 * CPU checkers export the same ABI under their own prefix (ref_* for oracle/_ref, orc_* for
 * oracle/port) and can never be mistaken for, or shadow, the product library's symbols. */
#ifndef ORACLE_PREFIX_H
#define ORACLE_PREFIX_H
#ifndef ORACLE_PREFIX
#error "define ORACLE_PREFIX (ref_ or orc_)"
#endif
#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC_NAME(n) ORC_CAT(ORACLE_PREFIX, n)
#define ptb_create             ORC_NAME(create)
#define ptb_destroy            ORC_NAME(destroy)
#define ptb_last_error         ORC_NAME(last_error)
#define ptb_version            ORC_NAME(version)
#define ptb_add_sphere         ORC_NAME(add_sphere)
#define ptb_add_plane          ORC_NAME(add_plane)
#define ptb_add_cylinder       ORC_NAME(add_cylinder)
#define ptb_add_pointset       ORC_NAME(add_pointset)
#define ptb_add_yarns          ORC_NAME(add_yarns)
#define ptb_add_mesh           ORC_NAME(add_mesh)
#define ptb_set_group_material ORC_NAME(set_group_material)
#define ptb_set_brdf           ORC_NAME(set_brdf)
#define ptb_add_merl           ORC_NAME(add_merl)
#define ptb_set_envmap         ORC_NAME(set_envmap)
#define ptb_set_light          ORC_NAME(set_light)
#define ptb_set_fog            ORC_NAME(set_fog)
#define ptb_set_background     ORC_NAME(set_background)
#define ptb_set_keyframes      ORC_NAME(set_keyframes)
#define ptb_set_frame          ORC_NAME(set_frame)
#define ptb_commit             ORC_NAME(commit)
#define ptb_render             ORC_NAME(render)
#define ptb_render_accum       ORC_NAME(render_accum)
#define ptb_resolve            ORC_NAME(resolve)
#define ptb_shard_pack_size    ORC_NAME(shard_pack_size)
#define ptb_shard_pack         ORC_NAME(shard_pack)
#define ptb_shard_unpack_add   ORC_NAME(shard_unpack_add)
#define ptb_primary_ids        ORC_NAME(primary_ids)
#define ptb_set_option         ORC_NAME(set_option)
#define ptb_get_scene_info     ORC_NAME(get_scene_info)
#define ptb_kat                ORC_NAME(kat)
#define ptb_get_kernel_times    ORC_NAME(get_kernel_times)
#define ptb_render_denoiser_inputs ORC_NAME(render_denoiser_inputs)
#define ptb_progressive_begin  ORC_NAME(progressive_begin)
#define ptb_progressive_pass   ORC_NAME(progressive_pass)
#define ptb_progressive_read   ORC_NAME(progressive_read)
/* options understood only by the CPU checkers */
#define ORC_OPT_THREADS 100    /* OpenMP threads for render (default: all, capped at 64 like the reference) */
#endif
