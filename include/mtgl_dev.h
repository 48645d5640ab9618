/*
 * mtgl_dev.h -- C ABI between the gl* front end and the sm_100a rendering back end.
 *
 * This is the drop-in boundary.  In the reference (zbufferoverflow/MyTinyGL) the same boundary is
 * the internal raster API of src/mytinygl.h:244-256 -- flush_points ... flush_polygon(GLState*),
 * transform_vertex, ndc_to_screen -- plus the places where src/gl_api.c touches the framebuffer
 * planes directly (glClear 409-457, glReadPixels 1180-1230, glDrawPixels 1286-1373) and the
 * object stores whose contents the rasteriser dereferences (textures.c:141-269, vbo.c:120-158).
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions: plain C, no C++/CUDA/torch types; every function returns 0 on success or a
 * negative MTGL_E_* code and never throws; the handle is opaque.  All calls are asynchronous with
 * respect to the GPU except mtgl_dev_finish / mtgl_dev_read_framebuffer.
 *
 * Two implementations of this header exist:
 *   - mytinygl_b200/csrc/dev/  : the product (hand-written CUDA for sm_100a).  No CPU fallback.
 *   - oracle/mtgl_oracle.c     : a scalar CPU restatement of the reference algorithm, linked only
 *                                into test binaries (test infrastructure, never shipped).
 */
#ifndef MTGL_DEV_H
#define MTGL_DEV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTGL_DEV_ABI_VERSION 6

/* error codes */
#define MTGL_OK            0
#define MTGL_E_NO_DEVICE  -1   /* no CUDA device / driver: the product never falls back to a CPU path */
#define MTGL_E_OOM        -2
#define MTGL_E_INVALID    -3
#define MTGL_E_CUDA       -4

/* capability bits of mtgl_state.caps (values follow src/mytinygl.h:26-37 so that a patched
 * gl_api.c can copy ctx->flags verbatim) */
#define MTGL_CAP_DEPTH_TEST     (1u << 1)
#define MTGL_CAP_CULL_FACE      (1u << 2)
#define MTGL_CAP_BLEND          (1u << 3)
#define MTGL_CAP_TEXTURE_2D     (1u << 4)
#define MTGL_CAP_LIGHTING       (1u << 5)
#define MTGL_CAP_FOG            (1u << 6)
#define MTGL_CAP_NORMALIZE      (1u << 7)
#define MTGL_CAP_COLOR_MATERIAL (1u << 8)
#define MTGL_CAP_ALPHA_TEST     (1u << 9)
#define MTGL_CAP_SCISSOR_TEST   (1u << 10)
#define MTGL_CAP_STENCIL_TEST   (1u << 11)

#define MTGL_MAX_LIGHTS 8

/* One light source: the fields of light_t (src/mytinygl.h:54-66) plus three quantities the
 * reference recomputes per vertex from state alone (src/lighting.h:77,101,103); the front end
 * evaluates them once with the same IEEE single-precision expression. */
typedef struct mtgl_light {
    float ambient[4];
    float diffuse[4];
    float specular[4];
    float position[4];        /* eye space; w == 0 -> directional */
    float dir_unit[3];        /* normalize(position.xyz)                 (lighting.h:77)  */
    float spot_exponent;
    float spot_dir_unit[3];   /* normalize(spot_direction)               (lighting.h:101) */
    float spot_cutoff;        /* degrees, 180 = not a spotlight */
    float cos_cutoff;         /* cosf(spot_cutoff * 3.14159265f / 180)   (lighting.h:103) */
    float att_constant;
    float att_linear;
    float att_quadratic;
    uint32_t enabled;
    uint32_t pad_[3];
} mtgl_light;

/* material_t (src/mytinygl.h:69-75) */
typedef struct mtgl_material {
    float ambient[4];
    float diffuse[4];
    float specular[4];
    float emission[4];
    float shininess;
    float pad_[3];
} mtgl_material;

/* Snapshot of every GLState field (src/mytinygl.h:86-225) the hot path reads.  Enumerated
 * fields hold the GL token itself (GL_LESS, GL_REPEAT, ...).  A vertex refers to the block that
 * was current when glVertex* was called (vertex stage, gl_api.c:263-348); a draw refers to the
 * block that was current at glEnd (raster stage, raster.c:451-725 reads ctx-> live). */
typedef struct mtgl_state {
    /* ---- vertex stage ---- */
    float modelview[16];      /* column major */
    float projection[16];
    float texture[16];
    float normal[12];         /* inverse-transpose 3x3 of modelview, columns padded to 4 (graphics.h:229-260) */
    mtgl_light lights[MTGL_MAX_LIGHTS];
    mtgl_material material_front;
    mtgl_material material_back;
    float light_model_ambient[4];
    uint32_t caps;                    /* MTGL_CAP_* */
    uint32_t light_model_local_viewer;
    uint32_t light_model_two_side;
    uint32_t color_material_face;     /* GL_FRONT / GL_BACK / GL_FRONT_AND_BACK */
    uint32_t color_material_mode;     /* GL_AMBIENT ... GL_AMBIENT_AND_DIFFUSE */
    uint32_t shade_model;             /* GL_FLAT / GL_SMOOTH / GL_PHONG */
    /* ---- clip / cull / setup ---- */
    int32_t  viewport[4];             /* x, y (top-down, raster.c:61-62), w, h */
    int32_t  scissor[4];
    uint32_t cull_face_mode;
    uint32_t front_face;
    uint32_t polygon_mode_front;
    uint32_t polygon_mode_back;
    /* ---- texture ---- */
    uint32_t texture_id;              /* GL texture name bound to GL_TEXTURE_2D (0 = none) */
    uint32_t tex_min_filter;
    uint32_t tex_mag_filter;
    uint32_t tex_wrap_s;
    uint32_t tex_wrap_t;
    uint32_t tex_env_mode;
    float    tex_env_color[4];
    uint32_t perspective_hint;
    /* ---- per-fragment ---- */
    uint32_t alpha_func;
    float    alpha_ref;
    uint32_t stencil_func;
    int32_t  stencil_ref;
    uint32_t stencil_mask;
    uint32_t stencil_fail;
    uint32_t stencil_zfail;
    uint32_t stencil_zpass;
    uint32_t stencil_writemask;
    uint32_t depth_func;
    uint32_t depth_mask;
    uint32_t blend_src;
    uint32_t blend_dst;
    uint32_t color_mask;              /* bit0 r, bit1 g, bit2 b, bit3 a */
    uint32_t fog_mode;
    float    fog_density;
    float    fog_start;
    float    fog_end;
    float    fog_color[4];
    float    line_width;
    float    point_size;
    uint32_t pad0_;
    double   depth_near;              /* GLdouble in the reference: the depth expression is evaluated */
    double   depth_far;               /* in double precision (raster.c:548) */
} mtgl_state;

/* An immediate-mode vertex as captured at glVertex* time: position (w is always 1,
 * gl_api.c:664-676) and the current colour / texcoord / normal, plus the state block index. */
typedef struct mtgl_in_vertex {
    float position[3];
    float color[4];
    float texcoord[2];
    float normal[3];
    uint32_t state;
} mtgl_in_vertex;

#define MTGL_TYPE_F32 0u
#define MTGL_TYPE_U8  1u

/* One enabled client array sourced from a buffer object (array_pointer_t, mytinygl.h:46-51,
 * resolved the way get_array_pointer/get_array_element do, gl_api.c:1745-1797). */
typedef struct mtgl_attrib {
    uint32_t enabled;
    uint32_t buffer;     /* glGenBuffers name */
    uint64_t offset;     /* byte offset of element 0 inside the buffer */
    uint32_t stride;     /* bytes between elements, already resolved (never 0) */
    uint16_t size;       /* components */
    uint16_t type;       /* MTGL_TYPE_* */
} mtgl_attrib;

#define MTGL_SRC_STAGED 0u   /* vertices are mtgl_in_vertex records in the batch          */
#define MTGL_SRC_ARRAYS 1u   /* vertices are fetched on the device from buffer objects    */

/* One glBegin/glEnd pair or one glDrawArrays/glDrawElements call. */
typedef struct mtgl_draw {
    uint32_t mode;           /* GL_POINTS ... GL_POLYGON */
    uint32_t count;          /* number of vertices */
    uint32_t raster_state;   /* mtgl_state index sampled at glEnd */
    uint32_t source;         /* MTGL_SRC_* */
    uint32_t first_staged;   /* STAGED: index of the first mtgl_in_vertex */
    uint32_t vertex_state;   /* ARRAYS: mtgl_state index for every vertex of the draw */
    int32_t  first;          /* ARRAYS: glDrawArrays 'first' (0 for glDrawElements) */
    uint32_t index_type;     /* ARRAYS: 0 = none, else GL_UNSIGNED_BYTE/SHORT/INT */
    uint32_t index_buffer;   /* ARRAYS: element buffer name, or 0 = indices live in batch.blob */
    uint32_t pad0_;
    uint64_t index_offset;   /* byte offset into the element buffer or the blob */
    mtgl_attrib position;
    mtgl_attrib color;
    mtgl_attrib texcoord;
    mtgl_attrib normal;
    float cur_color[4];      /* current values used where an array is disabled */
    float cur_texcoord[2];
    float cur_normal[3];
    uint32_t pad1_[3];
} mtgl_draw;

/* A batch: everything queued between two synchronisation points, in submission order.
 * If clear_mask != 0 the clear is executed before the first draw (glClear, gl_api.c:409-457:
 * honours the rectangle, ignores colour/depth/stencil write masks). */
typedef struct mtgl_batch {
    const mtgl_state     *states;    uint32_t n_states;    uint32_t pad0_;
    const mtgl_in_vertex *vertices;  uint32_t n_vertices;  uint32_t pad1_;
    const mtgl_draw      *draws;     uint32_t n_draws;     uint32_t pad2_;
    const void           *blob;      uint64_t blob_size;
    uint32_t clear_mask;             /* GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT | GL_STENCIL_BUFFER_BIT */
    int32_t  clear_rect[4];          /* x0, y0, x1, y1 (exclusive), already clamped to the framebuffer */
    uint32_t clear_color;            /* packed a<<24|b<<16|g<<8|r (graphics.h:337-348) */
    float    clear_depth;
    uint32_t clear_stencil;
} mtgl_batch;

/* Counters of the last submitted batch (device-side, valid after mtgl_dev_finish). */
typedef struct mtgl_dev_stats {
    uint64_t vertices;          /* vertices through the vertex stage */
    uint64_t triangles_in;      /* assembled triangles */
    uint64_t triangles_setup;   /* sub-triangles that survived clip / cull / degenerate tests */
    uint64_t tile_refs;         /* (sub-triangle, tile) pairs binned */
    uint64_t kernel_launches;   /* CUDA kernels launched by this library since creation */
    float    last_batch_ms;     /* CUDA-event time of the last batch (upload + all kernels) */
    float    stage_ms[5];       /* CUDA-event time per stage of the last batch, summed over passes:
                                   0 vertex (K1), 1 set-up (K2), 2 bin count + scan, 3 bin fill, 4 tile raster (K4/K5) */
    float    raster_ms[3];      /* stage 4 split by kernel group: 0 visibility (K4a), 1 shade (K4b), 2 general in-order kernel */
    /* the same times summed over every batch since creation (ABI v4): a caller that pipelines frames (glFlush per frame,
     * one glFinish at the end) reads these once instead of synchronising after each frame */
    uint64_t batches;
    double   cum_batch_ms;
    double   cum_stage_ms[5];
    double   cum_raster_ms[3];
    uint64_t chunks_culled;     /* 256-triangle chunks the culling pass dropped, last batch (k_cull.cu) */
} mtgl_dev_stats;

typedef struct mtgl_dev mtgl_dev;

/* gl_create_context -> framebuffer_init (gl_api.c:81-90, framebuffer.h:32-58): allocates the
 * colour (RGBA8), depth (f32) and stencil (u8) planes, row 0 = top, pitch = width.  'device' is a
 * CUDA ordinal (-1 = current device).  Planes are left uninitialised like the reference's. */
int mtgl_dev_create(int32_t width, int32_t height, int32_t device, mtgl_dev **out);
/* gl_destroy_context -> framebuffer_free (gl_api.c:238-250) */
void mtgl_dev_destroy(mtgl_dev *dev);

/* Sort-first band ownership for multi-GPU rendering: this device rasterises only framebuffer
 * rows [y0, y1).  Default is the whole framebuffer.  (No reference counterpart.) */
int mtgl_dev_set_band(mtgl_dev *dev, int32_t y0, int32_t y1);

/* Buffer ids: 1..256 are glGenBuffers names (vbo.h:17); ids MTGL_LIST_BUFFER_BASE + list (1..1024) belong to the front
 * end, which keeps the vertices of compiled display-list geometry there. */
#define MTGL_LIST_BUFFER_BASE 256u
#define MTGL_MAX_BUFFER_IDS   (257u + 1024u)

/* buffer_data / buffer_sub_data / buffer_delete (vbo.c:120-158, 96-110): device mirror of a
 * buffer object so that array draws fetch attributes on the device. */
int mtgl_dev_buffer_data(mtgl_dev *dev, uint32_t id, uint64_t size, const void *data);
int mtgl_dev_buffer_sub_data(mtgl_dev *dev, uint32_t id, uint64_t offset, uint64_t size, const void *data);
int mtgl_dev_buffer_delete(mtgl_dev *dev, uint32_t id);
/* read back part of a buffer-object mirror (the front end keeps no host copy of large buffers) */
/* device address and size of a buffer object's storage (NULL / 0 when it has none): lets a multi-GPU application fill
 * the buffer with a collective (each rank uploads a slice, NCCL all-gather over NVLink) instead of N full uploads */
int mtgl_dev_buffer_pointer(mtgl_dev *dev, uint32_t id, void **ptr, uint64_t *size);
int mtgl_dev_buffer_read(mtgl_dev *dev, uint32_t id, uint64_t offset, uint64_t size, void *out);

/* Pipelined transfers (no reference counterpart: the reference's buffers and framebuffer ARE host memory).
 *
 * mtgl_dev_buffer_data_pinned is mtgl_dev_buffer_data for a source in PAGE-LOCKED host memory that the caller leaves
 * unchanged until mtgl_dev_finish: the copy is queued on an upload stream and the call returns at once.  The buffer name
 * gets fresh storage (the storage that batches submitted earlier still read is orphaned and recycled once they have
 * finished -- what a GL driver does for glBufferData on a buffer in flight), so the upload of frame i+1 overlaps the
 * rasterisation and the read-back of frame i; batches submitted afterwards wait for the copy on the device.
 *
 * mtgl_dev_read_color_async queues a copy of colour rows [y0, y1) into page-locked host memory (`color` addresses row 0,
 * pitch = width) behind every batch submitted so far, on a read-back stream, and returns; batches submitted afterwards
 * do not touch the plane before the copy has left it.  mtgl_dev_finish waits for uploads, batches and read-backs. */
int mtgl_dev_buffer_data_pinned(mtgl_dev *dev, uint32_t id, uint64_t size, const void *pinned_data);
int mtgl_dev_read_color_async(mtgl_dev *dev, int32_t y0, int32_t y1, uint32_t *pinned_color);
/* mtgl_dev_buffer_orphan is the first half of mtgl_dev_buffer_data_pinned alone: the buffer name gets fresh storage of
 * `size` bytes that no submitted batch reads (contents undefined) and *ptr receives its device address.  The caller fills
 * it on a stream of its own -- a slice from host memory plus an NCCL all-gather over NVLink, say -- and orders the
 * context's stream (mtgl_dev_stream) behind that work with an event before it issues draws that read the buffer.  The
 * fill of frame i+1 then overlaps the rasterisation of frame i. */
int mtgl_dev_buffer_orphan(mtgl_dev *dev, uint32_t id, uint64_t size, void **ptr);

/* texture_upload_* (textures.c:141-269) after conversion to RGBA8 words (a<<24|b<<16|g<<8|r);
 * also builds mip level 1 the way texture_generate_mip1 does (textures.c:311-354). */
int mtgl_dev_texture_image(mtgl_dev *dev, uint32_t id, int32_t width, int32_t height, const uint32_t *rgba8);
int mtgl_dev_texture_delete(mtgl_dev *dev, uint32_t id);

/* flush_* (mytinygl.h:247-256) + emit_vertex (gl_api.c:263-348) + glClear (gl_api.c:409-457)
 * for a whole batch.  The call copies everything it needs before returning. */
int mtgl_dev_submit(mtgl_dev *dev, const mtgl_batch *batch);

/* glFinish (gl_api.c:895-899): block until all submitted batches have executed. */
int mtgl_dev_finish(mtgl_dev *dev);

/* Direct framebuffer plane access (ctx->framebuffer.*, glReadPixels gl_api.c:1180-1230 and
 * include/mytinygl/sdl.h:76-81 read; glDrawPixels gl_api.c:1286-1373 writes).  NULL = skip a plane.
 * Rows [y0, y1) of each plane are copied to/from host arrays laid out like the reference's
 * (full-frame pitch = width; the pointer addresses row 0). */
int mtgl_dev_read_framebuffer(mtgl_dev *dev, int32_t y0, int32_t y1,
                              uint32_t *color, float *depth, uint8_t *stencil);
int mtgl_dev_write_framebuffer(mtgl_dev *dev, int32_t y0, int32_t y1,
                               const uint32_t *color, const float *depth, const uint8_t *stencil);

/* Device addresses of the planes (for the multi-GPU gather and for zero-copy consumers). */
int mtgl_dev_plane_pointers(mtgl_dev *dev, void **color, void **depth, void **stencil);

/* Multi-GPU present path (no reference counterpart: the reference has one framebuffer in host memory that mtgl_swap
 * reads, include/mytinygl/sdl.h:76-81).  The presenting rank exports its colour plane as a CUDA IPC handle
 * (MTGL_IPC_HANDLE_BYTES opaque bytes, to be carried to the other processes by any means); a rank that is given the
 * handle maps the plane over NVLink and every colour store of its band -- shade kernel, tile write-back, clears,
 * mtgl_dev_write_framebuffer -- is mirrored into it, so that after all ranks have finished their frame the presenting
 * rank's plane is complete without a separate gather.  Passing NULL detaches. */
#define MTGL_IPC_HANDLE_BYTES 64
int mtgl_dev_export_color_plane(mtgl_dev *dev, void *handle_out);
int mtgl_dev_set_present_target(mtgl_dev *dev, const void *handle);
/* How a band reaches the presenting GPU's plane.
 *   MTGL_PRESENT_STORES (default): fused -- every raster kernel stores its finished tiles locally AND into the mapped plane.
 *   MTGL_PRESENT_COPY: the kernels store locally only; mtgl_dev_frame_barrier pushes the band with one asynchronous
 *     peer-to-peer copy over NVLink on a side stream, followed there by the barrier.  The next frame's geometry and
 *     visibility stages (which do not touch the colour plane) run meanwhile; only its colour-writing kernels wait.  This is
 *     the better choice when the presenting GPU's NVLink ingest is the bottleneck (8 GPUs sending an 8K frame: 116 MB per
 *     frame into one GPU), where fused stores stall the SMs that issue them. */
#define MTGL_PRESENT_STORES 0
#define MTGL_PRESENT_COPY   1
int mtgl_dev_set_present_mode(mtgl_dev *dev, int mode);

/* Pixel rectangles.  mtgl_dev_draw_pixels replaces the loop of glDrawPixels (gl_api.c:1286-1373):
 * the rectangle is copied at call time and drawn in stream order behind every batch submitted so far -- no host
 * synchronisation.  'caps' holds MTGL_CAP_ALPHA_TEST / MTGL_CAP_DEPTH_TEST / MTGL_CAP_BLEND as sampled at the call;
 * formats: GL_RGBA, GL_RGB, GL_LUMINANCE, GL_LUMINANCE_ALPHA of unsigned bytes, rows tightly packed, first row = bottom.
 * (x, y) is the raster position in GL window coordinates (origin bottom-left).  Rows outside the device's band are
 * skipped.  mtgl_dev_read_pixels replaces the loop of glReadPixels (1180-1230) for GL_RGBA / GL_RGB: waits for the
 * device, then returns width x height x bpp bytes (zeros outside the framebuffer rows, (0,0,0,255) outside its columns). */
typedef struct mtgl_pixel_rect {
    int32_t  x, y, width, height;
    uint32_t format;
    uint32_t caps;
    uint32_t alpha_func, depth_func;   /* GL tokens */
    float    alpha_ref;
    uint32_t depth_mask;
    uint32_t blend_src, blend_dst;     /* GL tokens */
} mtgl_pixel_rect;
int mtgl_dev_draw_pixels(mtgl_dev *dev, const mtgl_pixel_rect *rect, const void *pixels);
int mtgl_dev_read_pixels(mtgl_dev *dev, int32_t x, int32_t y, int32_t width, int32_t height, uint32_t format, void *out);

/* Frame barrier of a multi-GPU frame, queued on this context's stream behind its raster kernels: returns at once; the
 * stream continues when all 'participants' contexts (the presenter, i.e. the one that exported its plane, and every
 * context that mapped it) have queued the same barrier and finished the work in front of it.  The counter lives behind
 * the presenter's colour plane, so nothing but the exported handle is needed.  Every participant must call it the same
 * number of times after export / set_present_target; participants == 1 is a no-op.  (No reference counterpart.) */
int mtgl_dev_frame_barrier(mtgl_dev *dev, uint32_t participants);

int mtgl_dev_get_stats(mtgl_dev *dev, mtgl_dev_stats *out);

/* Device-side stopwatch on the context's stream: mark(0) ... mark(1), then elapsed = CUDA-event time
 * between the two marks (waits for mark 1).  Used by bench.py to time K steps on the device. */
int mtgl_dev_timer_mark(mtgl_dev *dev, int which);
/* the CUDA stream (cudaStream_t) all of this context's work is queued on: lets a multi-GPU application order its own
 * collectives against the frames with events instead of host synchronisation (bench.py) */
void *mtgl_dev_stream(mtgl_dev *dev);
int mtgl_dev_timer_elapsed_ms(mtgl_dev *dev, float *ms);
const char *mtgl_dev_last_error(mtgl_dev *dev);
int mtgl_dev_abi_version(void);

#ifdef __cplusplus
}
#endif

#endif /* MTGL_DEV_H */
