/*
 * gl_state.h -- host-side GL state of the B200-native MyTinyGL front end.
 *
 * Field-for-field counterpart of the reference's GLState (src/mytinygl.h:86-225) with two
 * differences: the framebuffer lives in HBM behind include/mtgl_dev.h (the host keeps a lazily
 * synchronised mirror), and draw calls are queued into a batch instead of being rasterised
 * synchronously at glEnd.
 */
#ifndef MTGL_FRONT_GL_STATE_H
#define MTGL_FRONT_GL_STATE_H

#include <cstdint>
#include <vector>

#include "GL/gl.h"
#include "mtgl_context.h"
#include "mtgl_dev.h"

namespace mtgl {

constexpr int kMatrixStackDepth = 24;   /* mytinygl.h:21 */
constexpr int kMaxLights = 8;           /* mytinygl.h:22 */
constexpr int kMaxListDepth = 64;       /* mytinygl.h:23 */
constexpr int kMaxTextures = 256;       /* textures.h:18 */
constexpr int kMaxBuffers = 256;        /* vbo.h:17 */
constexpr int kMaxLists = 1024;         /* lists.h:18 */
constexpr int kMaxTextureSize = 2048;   /* textures.h:17 */

struct Rgba { float r, g, b, a; };

struct Light {                          /* light_t, mytinygl.h:54-66 */
    Rgba ambient, diffuse, specular;
    float position[4];
    float spot_direction[3];
    float spot_exponent, spot_cutoff;
    float att_constant, att_linear, att_quadratic;
    GLboolean enabled;
};

struct Material {                       /* material_t, mytinygl.h:69-75 */
    Rgba ambient, diffuse, specular, emission;
    float shininess;
};

struct ArrayPointer {                   /* array_pointer_t, mytinygl.h:46-51 */
    GLint size;
    GLenum type;
    GLsizei stride;
    const void *pointer;
};

struct Texture {                        /* texture_t, textures.h:21-39 (level 1 lives on the device) */
    bool allocated = false;
    int32_t width = 0, height = 0;
    std::vector<uint32_t> pixels;
    GLint min_filter = GL_NEAREST, mag_filter = GL_NEAREST, wrap_s = GL_REPEAT, wrap_t = GL_REPEAT;
};

struct Buffer {                         /* buffer_t, vbo.h */
    bool allocated = false;
    bool has_data = false;              /* storage exists (glBufferData with size > 0) */
    bool host_valid = false;            /* 'data' mirrors the contents; large buffers live in HBM only */
    uint64_t size = 0;
    std::vector<uint8_t> data;
    GLenum usage = 0;
    /* HBM-only buffers: the few bytes the host has looked at (the last element of an array draw becomes the current
     * normal / colour / texture coordinate, gl_api.c:1826-1842).  Kept up to date from the caller's memory by
     * glBufferData / glBufferSubData, so a frame loop never waits for the device to learn them again. */
    struct Peek { uint64_t off; uint32_t n; uint8_t raw[16]; };
    std::vector<Peek> peeks;
    bool peeks_frozen = false;          /* the storage's address was handed out (mtgl_context_buffer_pointer): nothing about its
                                         * contents may be remembered until the next glBufferData */
};

constexpr uint64_t kHostMirrorLimit = 8u << 20;   /* buffers above this keep no host copy */

enum ListOp : uint8_t {                 /* the 31 opcodes of lists.h:21-53 */
    OP_END, OP_BEGIN, OP_VERTEX, OP_COLOR, OP_TEXCOORD, OP_NORMAL, OP_TRANSLATE, OP_ROTATE, OP_SCALE,
    OP_PUSH_MATRIX, OP_POP_MATRIX, OP_LOAD_IDENTITY, OP_MULT_MATRIX, OP_LOAD_MATRIX, OP_MATRIX_MODE,
    OP_ORTHO, OP_FRUSTUM, OP_ENABLE, OP_DISABLE, OP_BIND_TEXTURE, OP_BLEND_FUNC, OP_DEPTH_FUNC,
    OP_DEPTH_MASK, OP_CULL_FACE, OP_FRONT_FACE, OP_SHADE_MODEL, OP_LIGHTF, OP_LIGHTFV, OP_MATERIALF,
    OP_MATERIALFV, OP_CALL_LIST
};

struct ListCmd {
    ListOp op;
    GLenum e[2];          /* enum / integer operands */
    union {
        float f[16];
        double d[6];
    };
};

/* A glBegin ... glEnd stretch of a display list that holds nothing but vertices and their attributes, compiled at
 * glEndList into one array draw from a device buffer the list owns (12 floats per vertex: position, colour, texture
 * coordinate, normal): glCallList then queues one draw record instead of replaying every glVertex / glNormal call. */
struct ListRun {
    size_t begin_cmd = 0, end_cmd = 0;      /* indices of OP_BEGIN and OP_END in cmds */
    GLenum mode = 0;
    uint32_t first = 0, count = 0;          /* vertices in the list's buffer */
    bool has_color = false, has_texcoord = false, has_normal = false;   /* set inside the run, before its first vertex */
    float last_color[4], last_texcoord[2], last_normal[3];              /* what the run leaves as current values */
};

struct DisplayList {
    bool allocated = false;
    bool valid = false;
    std::vector<ListCmd> cmds;
    std::vector<ListRun> runs;              /* sorted by begin_cmd */
    bool has_buffer = false;                /* device buffer MTGL_LIST_BUFFER_BASE + list id holds the runs' vertices */
};

constexpr uint32_t kListRunMinVertices = 24;    /* shorter runs are replayed: a draw record costs more than they do */

} // namespace mtgl

struct GLState {
    /* clear values */
    mtgl::Rgba clear_color;
    GLdouble clear_depth;
    GLint stencil_clear;

    GLint viewport_x, viewport_y;
    GLsizei viewport_w, viewport_h;

    mtgl::Rgba current_color;
    float current_texcoord[2];
    float current_normal[3];

    GLenum matrix_mode;
    GLfloat modelview[mtgl::kMatrixStackDepth][16];
    GLfloat projection[mtgl::kMatrixStackDepth][16];
    GLfloat texture[mtgl::kMatrixStackDepth][16];
    GLint modelview_depth, projection_depth, texture_depth;

    GLenum primitive_mode;
    bool inside_begin_end;
    uint32_t caps;                      /* MTGL_CAP_* */

    GLenum blend_src, blend_dst;
    GLenum cull_face_mode, front_face;
    GLenum depth_func;
    GLboolean depth_mask;
    GLenum alpha_func;
    GLfloat alpha_ref;
    GLint scissor_x, scissor_y;
    GLsizei scissor_w, scissor_h;
    GLenum stencil_func;
    GLint stencil_ref;
    GLuint stencil_mask;
    GLenum stencil_fail, stencil_zfail, stencil_zpass;
    GLuint stencil_writemask;
    GLdouble depth_near, depth_far;
    GLboolean color_mask[4];
    GLfloat line_width, point_size;
    GLenum polygon_mode_front, polygon_mode_back;

    std::vector<mtgl::Texture> textures;
    GLuint bound_texture_2d;
    GLenum tex_env_mode;
    mtgl::Rgba tex_env_color;
    GLenum perspective_hint;

    GLint raster_pos_x, raster_pos_y;
    GLboolean raster_pos_valid;

    GLenum fog_mode;
    GLfloat fog_density, fog_start, fog_end;
    mtgl::Rgba fog_color;

    mtgl::Light lights[mtgl::kMaxLights];
    mtgl::Material material_front, material_back;
    mtgl::Rgba light_model_ambient;
    GLboolean light_model_local_viewer, light_model_two_side;
    GLenum color_material_face, color_material_mode;
    GLenum shade_model;

    std::vector<mtgl::Buffer> buffers;
    GLuint bound_array_buffer, bound_element_buffer;

    uint32_t client_state;              /* bit0 vertex, bit1 colour, bit2 texcoord, bit3 normal */
    mtgl::ArrayPointer vertex_pointer, color_pointer, texcoord_pointer, normal_pointer;

    std::vector<mtgl::DisplayList> lists;
    GLuint list_base, list_index;
    GLenum list_mode;
    GLuint list_call_depth;
    uint64_t list_runs_drawn = 0;           /* compiled ListRuns queued as array draws (statistics) */

    GLenum error;

    /* ---- device + batching (no reference counterpart) ---- */
    mtgl_dev *dev;
    int32_t fb_width, fb_height;
    mtgl_framebuffer mirror;            /* host mirror of the planes */
    std::vector<uint32_t> mirror_color;
    std::vector<float> mirror_depth;
    std::vector<uint8_t> mirror_stencil;

    std::vector<mtgl_state> states;
    std::vector<mtgl_in_vertex> staged;
    std::vector<mtgl_draw> draws;
    std::vector<uint8_t> blob;
    uint32_t prim_first;                /* first staged vertex of the primitive being assembled */
    bool vstate_dirty;                  /* a state-changing call happened since the last snapshot */
    bool vstate_full_dirty;             /* ... and it was not just a matrix call: the snapshot has to be rebuilt from scratch */
    bool have_snapshot;                 /* last_snapshot mirrors the live state up to the current matrices */
    mtgl_state last_snapshot;
    bool material_touched;              /* COLOR_MATERIAL rewrote ctx materials since the last snapshot */
    uint32_t vstate_index;              /* state block of the most recent vertex */
    uint32_t pending_clear_mask;
    int32_t pending_clear_rect[4];
    uint32_t pending_clear_color;
    float pending_clear_depth;
    uint32_t pending_clear_stencil;
};

#endif
