/*
 * swr_b200.h -- C ABI of the B200-native SoftwareRenderer draw path (libswr_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree,
 * src/renderer/...).  The C++ classes in include/swr/Renderer.h (same names and signatures as
 * the reference's VertexProcessor / Rasterizer / shader bases) are thin inline shims over
 * these calls; a Python / ctypes binding is in softwarerenderer_b200/api.py; INTEGRATION.md
 * shows the binding a maintainer of the reference would add.
 *
 * Execution model: a context owns one CUDA stream and its device scratch.  Draws are
 * enqueued asynchronously on that stream; swr_finish() waits and reports deferred errors.
 * There is no CPU fallback: every draw runs the sm_100a kernels or fails.
 *
 * All functions returning int return 0 on success and a negative value on error;
 * swr_last_error() then describes it.
 */
#ifndef SWR_B200_H
#define SWR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SWR_API
#else
#define SWR_API __attribute__((visibility("default")))
#endif

/* ---- enums: same values as the reference's enum classes ---------------------------------- */
enum { SWR_DRAW_POINT = 0, SWR_DRAW_LINE = 1, SWR_DRAW_TRIANGLE = 2 };          /* VertexProcessor.h:42-46 */
enum { SWR_CULL_NONE = 0, SWR_CULL_CCW = 1, SWR_CULL_CW = 2 };                 /* VertexProcessor.h:49-53 */
enum { SWR_RASTER_SPAN = 0, SWR_RASTER_BLOCK = 1, SWR_RASTER_ADAPTIVE = 2 };    /* Rasterizer.h:45-49 */

/* ---- constants ------------------------------------------------------------------------------ */
#define SWR_BLOCK_SIZE 8            /* IRasterizer.h:33 */
#define SWR_MAX_AVARS 16            /* IRasterizer.h:36 */
#define SWR_MAX_PVARS 16            /* IRasterizer.h:39 */
#define SWR_MAX_VERTEX_ATTRIBS 8    /* VertexConfig.h:34 */
#define SWR_BATCH_PRIMS 1024        /* VertexProcessor.cpp:110 */
#define SWR_MAX_POLY 12             /* clipped polygon cap (9 in general position; see oracle/swr_scene.h) */
#define SWR_ORDINAL_STRIDE 10240u   /* emission ordinal = batch * stride + slot */
#define SWR_MAX_TILE_MIRRORS 7
#define SWR_MAX_RANKS 8            /* GPUs of one sort-first partition (one NVLink / NVSwitch domain) */
#define SWR_MAX_RENDER_TARGETS 12
#define SWR_MAX_UNIFORM_BYTES 1024

/* ---- stock shader pack (compiled into libswr_b200.so; ids shared with oracle/swr_scene.h) - */
enum { SWR_VS_POS_COLOR = 0, SWR_VS_MVP_COLOR = 1, SWR_VS_MVP_NORMAL_UV = 2, SWR_VS_COUNT = 3 };
enum { SWR_PS_FLAT = 0, SWR_PS_COUNT_ID = 1, SWR_PS_GOURAUD = 2, SWR_PS_GOURAUD_DEPTH = 3,
       SWR_PS_VARY_DUMP = 4, SWR_PS_TEXTURED = 5, SWR_PS_TEXTURED_ANISO = 6, SWR_PS_COUNT = 7 };

/* Render-target slots used by the stock pixel shaders. */
enum { SWR_RT_COLOR = 0, SWR_RT_DEPTH = 1, SWR_RT_COUNT = 2, SWR_RT_PRIM_ID = 3, SWR_RT_VARY0 = 4 /* ..11 */ };

/* Uniform block read by the stock shaders (swr_set_uniforms). */
#define SWR_MAX_MIP_LEVELS 14
typedef struct swr_stock_uniforms {
    float mvp[16];               /* row-major: clip.x = m[0]*x + m[1]*y + m[2]*z + m[3]*w */
    const uint32_t *texture;     /* device pointer, tex_w * tex_h texels (0x00RRGGBB) */
    int32_t tex_w, tex_h;        /* powers of two */
    /* mip chain for SWR_PS_TEXTURED_ANISO (same layout as swr::TextureView, include/swr/Texture.h) */
    const uint32_t *mip[SWR_MAX_MIP_LEVELS];
    int32_t mip_w[SWR_MAX_MIP_LEVELS];
    int32_t mip_h[SWR_MAX_MIP_LEVELS];
    int32_t mip_levels;
    int32_t max_anisotropy;
} swr_stock_uniforms;

/* ---- shader programs ------------------------------------------------------------------------ */
/* What VertexProcessor::setVertexShader<VS>() captures (VertexProcessor.h:77-85): instead of a
 * pointer to VS::processVertex, a host launcher of the geometry kernel instantiated with VS. */
typedef void (*swr_launch_fn)(const void *args, void *cuda_stream);
typedef int (*swr_uniform_fn)(const void *data, size_t bytes, void *cuda_stream);

typedef struct swr_vertex_shader {
    swr_launch_fn launch_geometry;   /* args: swr::detail::GeomArgs */
    swr_launch_fn launch_stream_out; /* the vertex stage alone, into RasterizerVertex / index arrays (swr_process_elements) */
    swr_uniform_fn set_uniforms;     /* copies the uniform block into the shader TU's __constant__ memory */
    int32_t attrib_count, avar_count, pvar_count;
    const char *name;
    uint32_t args_layout;            /* SWR_ARGS_LAYOUT of the headers the shader TU was compiled with */
} swr_vertex_shader;

/* What Rasterizer::setPixelShader<PS>() captures (Rasterizer.h:90-96): launchers of the tile
 * kernel instantiated with PS, one per draw mode and tile size. */
typedef struct swr_pixel_shader {
    swr_launch_fn launch_tiles[3][2];   /* [draw mode][tile size index: 0 = 32 px, 1 = 64 px] */
    swr_uniform_fn set_uniforms;
    int32_t avar_count, pvar_count, interpolate_z, interpolate_w;
    int32_t render_targets;             /* slots [0, render_targets) are staged in shared memory */
    const char *name;
    uint32_t args_layout;               /* SWR_ARGS_LAYOUT of the headers the shader TU was compiled with */
} swr_pixel_shader;

SWR_API const swr_vertex_shader *swr_stock_vertex_shader(int vs_kind);
SWR_API const swr_pixel_shader *swr_stock_pixel_shader(int ps_kind);

/* ---- context -------------------------------------------------------------------------------- */
typedef struct swr_context swr_context;

SWR_API int swr_create(swr_context **out, int cuda_device);
SWR_API void swr_destroy(swr_context *ctx);
SWR_API const char *swr_last_error(void);
SWR_API int swr_abi_version(void);

/* ---- VertexProcessor state (VertexProcessor.h:59-91, VertexProcessor.cpp:29-76) ------------- */
SWR_API int swr_set_viewport(swr_context *ctx, int x, int y, int width, int height);      /* setViewport   */
SWR_API int swr_set_depth_range(swr_context *ctx, float n, float f);                      /* setDepthRange */
SWR_API int swr_set_cull_mode(swr_context *ctx, int cull_mode);                           /* setCullMode   */
/* setVertexAttribPointer(index, stride, buffer).  `bytes` is the additive extent: required (> 0)
 * when `buffer` is host memory (it is staged to the device per draw), ignored for device memory. */
SWR_API int swr_set_vertex_attrib_pointer(swr_context *ctx, int index, int stride, const void *buffer, size_t bytes);
SWR_API int swr_set_vertex_shader(swr_context *ctx, const swr_vertex_shader *vs);         /* setVertexShader<VS> */

/* ---- Rasterizer state (Rasterizer.h:67-96) -------------------------------------------------- */
SWR_API int swr_set_raster_mode(swr_context *ctx, int raster_mode);                       /* setRasterMode  */
SWR_API int swr_set_scissor_rect(swr_context *ctx, int x, int y, int width, int height);  /* setScissorRect */
SWR_API int swr_set_pixel_shader(swr_context *ctx, const swr_pixel_shader *ps);           /* setPixelShader<PS> */

/* ---- additive surface: render targets, uniforms, tiling ------------------------------------- */
/* Register a 32-bit-per-pixel device surface.  The tile kernel stages the registered slots the
 * pixel shader declares in shared memory and writes them back with 128-bit stores; all slots
 * share width/height (the tile grid).  ptr == NULL unregisters. */
SWR_API int swr_set_render_target(swr_context *ctx, int slot, void *device_ptr, int pitch_bytes, int width, int height);
SWR_API int swr_set_uniforms(swr_context *ctx, const void *data, size_t bytes);
/* tile_size: 0 = automatic, 32 or 64.  Sort-first partition: this context rasterizes only the
 * screen tiles t with (tx + 3*ty) % world == rank (geometry is processed in full). */
SWR_API int swr_set_tile_size(swr_context *ctx, int tile_size);
SWR_API int swr_set_tile_partition(swr_context *ctx, int rank, int world);
/* Host index arrays of big draws (>= 32 MB, uploaded slice by slice under the kernels): a slice whose blocks of 4096
 * consecutive indices each span at most 65535 vertices is sent as 16-bit offsets + one 32-bit base per block (packed by
 * a few host threads while the previous slice is on the link) and widened on the device -- lossless, half the bytes.
 * mode -1 = automatic (kept while this host packs faster than ~32 GB/s, i.e. faster than PCIe would have carried the
 * saved bytes), 0 = off, 1 = always try.  The reference passes the same int32 array (VertexProcessor.h:90). */
SWR_API int swr_set_index_narrowing(swr_context *ctx, int mode);
/* test hook (no GPU needed): the host-side packer; returns 1 = packed, 0 = some block spans more than 16 bits,
 * -1 = this CPU has no AVX2 (narrowing is then never used).  bases needs ceil(count / 4096) entries. */
SWR_API int swr_debug_pack_indices16(const int32_t *indices, size_t count, uint16_t *offsets, int32_t *bases);
/* Heavy-tile split: a tile whose binning pass lists more than `groups` 32-record groups is shaded by four CTAs, one
 * per quadrant (results are unchanged: every pixel still sees its fragments in emission order).  -1 = automatic
 * (currently off: on the measured meshes the repeated record tests cost more than the shorter tail saves), 0 = off.
 * No reference counterpart (scheduling only). */
SWR_API int swr_set_tile_split(swr_context *ctx, int groups);
SWR_API int swr_set_scratch_limit(swr_context *ctx, size_t bytes);
/* Enqueue on a caller-owned CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)
 * instead of the context's own; NULL restores the context's stream.  Waits for pending work first. */
SWR_API int swr_set_stream(swr_context *ctx, void *cuda_stream);
/* Stage overlap (default off).  enable: the geometry kernel runs on an auxiliary stream with two alternating
 * scratch sets, so the geometry of pass k+1 can run under the tile kernel of pass k; passes_hint > 1 cuts a draw
 * into that many passes.  overlap_draws: the geometry of the NEXT draw may also start while the tiles of the
 * previous draw run; the caller then must not modify vertex / index buffers or uniforms between draws without
 * swr_finish.  Measured on B200: the tile kernel fills the SMs' registers, so the gain is 2-10 %
 * across draws and negative for multi-pass draws (every tile pass pays its own tail). */
SWR_API int swr_set_pipeline(swr_context *ctx, int enable, int passes_hint, int overlap_draws);

/* ---- draws ---------------------------------------------------------------------------------- */
/* VertexProcessor::drawElements(mode, count, indices) (VertexProcessor.cpp:78-120).  `indices`
 * may be host or device memory.  Asynchronous. */
SWR_API int swr_draw_elements(swr_context *ctx, int draw_mode, size_t count, const int32_t *indices);
/* IRasterizer::draw{Point,Line,Triangle}List (IRasterizer.h:56-71, Rasterizer.h:116-141) on
 * screen-space RasterizerVertex records (144 bytes each, IRasterizer.h:42-53), host or device. */
SWR_API int swr_draw_raster_list(swr_context *ctx, int draw_mode, const void *vertices, size_t vertex_count,
                                 const int32_t *indices, size_t index_count);
SWR_API int swr_finish(swr_context *ctx);
/* VertexProcessor::drawElements for a VertexProcessor whose rasterizer is a user's IRasterizer (IRasterizer.h:56-71,
 * VertexProcessor.cpp:302-317): runs the vertex stage only (vertex shader, clipping, perspective divide, viewport,
 * culling) and calls `emit` once per batch of 1024 input primitives, in order, with exactly what the reference hands
 * to IRasterizer::draw{Point,Line,Triangle}List: screen-space RasterizerVertex records (144 bytes each) and the index
 * list -- -1 for dropped primitives, re-oriented triangles with their first and last index swapped, the clipper's fan
 * triangles appended.  Host arrays, valid during the call only.  Synchronous. */
typedef void (*swr_stream_out_fn)(void *user, int draw_mode, const void *vertices, size_t vertex_count,
                                  const int32_t *indices, size_t index_count);
SWR_API int swr_process_elements(swr_context *ctx, int draw_mode, size_t count, const int32_t *indices,
                                 swr_stream_out_fn emit, void *user);

/* ---- measurement ---------------------------------------------------------------------------- */
typedef struct swr_stats {
    uint64_t fragments;          /* drawPixel invocations since swr_reset_stats */
    uint64_t primitives_in;      /* primitives submitted */
    uint64_t kernel_launches;    /* kernels of this library launched */
    uint64_t draws;
    uint64_t passes;
    float last_geometry_ms;      /* device time of the last draw: first to last geometry launch / first tile-phase launch to */
    float last_tile_ms;          /*   the end of the draw (CUDA events on the context stream; valid after swr_finish).  In a   */
                                 /*   multi-pass draw the passes interleave, so the two intervals overlap and do not add up.   */
    int32_t last_tile_size;
    int32_t reserved;
    uint64_t scratch_bytes;
    uint64_t h2d_bytes;          /* bytes the library's staging copied host -> device (vertex attributes, indices) */
} swr_stats;
SWR_API int swr_get_stats(swr_context *ctx, swr_stats *out);   /* synchronizes the stream */
SWR_API int swr_reset_stats(swr_context *ctx);
SWR_API int swr_timer_begin(swr_context *ctx);                 /* CUDA event on the context stream */
SWR_API int swr_timer_end(swr_context *ctx, float *ms);        /* records, synchronizes, returns elapsed */

/* ---- device memory helpers for FFI hosts without their own CUDA binding --------------------- */
SWR_API void *swr_device_alloc(swr_context *ctx, size_t bytes);
SWR_API int swr_device_free(swr_context *ctx, void *ptr);
SWR_API void *swr_host_alloc_pinned(size_t bytes);
SWR_API int swr_host_free_pinned(void *ptr);
SWR_API int swr_memcpy_h2d(swr_context *ctx, void *dst, const void *src, size_t bytes);   /* async on the stream */
SWR_API int swr_memcpy_d2h(swr_context *ctx, void *dst, const void *src, size_t bytes);   /* async on the stream */
SWR_API int swr_memset32(swr_context *ctx, void *dst, uint32_t value, size_t count);      /* async on the stream */
SWR_API int swr_flush_l2(swr_context *ctx);   /* overwrites a 256 MiB scratch buffer (bench hygiene) */

/* ---- multi-GPU composite helpers (sort-first tiles; the exchange itself is NCCL, done by the host) */
/* Copies the tiles owned by (rank, world) of a registered slot into a dense tile-major buffer
 * (owned tiles in increasing tile id, TILE*TILE words each) and back. */
SWR_API int swr_pack_tiles(swr_context *ctx, int slot, int rank, int world, int tile_size, void *dst_device);
SWR_API int swr_unpack_tiles(swr_context *ctx, int slot, int rank, int world, int tile_size, const void *src_device);
SWR_API int64_t swr_owned_tile_count(int width, int height, int tile_size, int rank, int world);

/* Composite fused into the tile kernel: every finished tile of render-target slot `slot` is also stored to
 * `count` (<= SWR_MAX_TILE_MIRRORS) other surfaces of the same pitch and size -- the peers' framebuffers, mapped
 * with swr_ipc_open, so the stores travel over NVLink while the remaining tiles are still being shaded and no
 * pack / all-gather / unpack pass is needed.  Tiles that no primitive touched are not stored (every rank is
 * expected to clear its surface the same way); the host provides the cross-rank barriers: one after the draws
 * (before a mirror is read) and one between clearing a mirrored surface and the first draw of the next frame.
 * count = 0 switches it off. */
SWR_API int swr_set_tile_mirrors(swr_context *ctx, int slot, int count, void *const *surfaces);
/* CUDA IPC plumbing for the above: handle (64 bytes) + offset of a device pointer inside its allocation,
 * and the mapping of such a handle into this process / device. */
SWR_API int swr_ipc_get_handle(const void *device_ptr, void *handle64, int64_t *offset);
SWR_API int swr_ipc_open(swr_context *ctx, const void *handle64, int64_t offset, void **device_ptr);
SWR_API int swr_ipc_close(swr_context *ctx, void *device_ptr);

/* ---- sharded geometry: the vertex stage of a sort-first partition, split by batches ------------------------
 * With swr_set_tile_partition alone every rank runs the whole vertex stage and keeps the records of its own tiles.
 * With geometry shards each rank runs only every world-th run of 16 batches and writes every surviving record
 * straight into the scratch of the ranks whose tiles it touches -- plain stores through peer mappings (NVLink), in
 * the reference's emission order per destination -- followed by a flag barrier between the GPUs (no NCCL call on
 * the path).  Set-up, identical on every rank: (1) swr_shared_scratch_create(bytes): one allocation that holds the
 * per-pass scratch (same `bytes`, render-target size, tile size and scratch limit on every rank: the ranks address
 * each other's arrays by offset); (2) export it with swr_ipc_get_handle, exchange the handles, map the peers' with
 * swr_ipc_open; (3) a host barrier (all arenas exist and are zeroed); (4) swr_set_geometry_shards(rank, world,
 * arenas) with arenas[r] = rank r's scratch as mapped HERE (own entry ignored).  Every rank must then issue the same
 * sequence of draws and swr_peer_barrier calls.  world = 1 switches it off. */
SWR_API int swr_shared_scratch_create(swr_context *ctx, size_t bytes, void **base);
SWR_API int swr_set_geometry_shards(swr_context *ctx, int rank, int world, void *const *arenas);
/* Stream-ordered barrier between the ranks of the partition (one 32-thread kernel: release stores to the peers'
 * flag words, acquire spin on the own ones).  What every rank enqueued before it -- draws, tile mirror stores -- is
 * complete and visible on all ranks before anything enqueued after it starts.  The frame loop of a partition calls
 * it once per frame (composite complete); it also pre-pays the barrier the next draw would otherwise need. */
SWR_API int swr_peer_barrier(swr_context *ctx);

/* ---- debugging ------------------------------------------------------------------------------ */
/* Per-tile timing of the next draws: after a draw, swr_debug_read_tile_stats copies 16 uint32 per tile
 * {start ns (low 32 bits of %globaltimer), duration ns, primitives queued, fragments, clocks/16 of thread 0 in the
 * pre-test / coverage / shading phases, flushes, clocks/16 in the flush prologue / record binning / chunk + group
 * binning, records tested, groups tested, 0, 0, 0}; returns the tile count.  Only in builds with
 * -DSWR_TILE_STATS=1 (the clocks cost ~2 %); the stock build returns an error from the enable call. */
SWR_API int swr_debug_enable_tile_stats(swr_context *ctx, int enable);
SWR_API int64_t swr_debug_read_tile_stats(swr_context *ctx, uint32_t *out, int64_t cap_tiles);

#ifdef __cplusplus
}
#endif
#endif /* SWR_B200_H */
