/*
 * GL/gl.h -- public OpenGL 1.5 fixed-function API of the B200-native MyTinyGL back end.
 *
 * Drop-in for the reference's include/GL/gl.h (zbufferoverflow/MyTinyGL, include/GL/gl.h:18-669):
 * same C linkage, same scalar typedefs, same token values (Khronos registry numbering plus the
 * reference's private GL_PHONG = 0x1d02) and the same 111 entry points.  The entry points are
 * implemented by mytinygl_b200/csrc/front/ and forward rendering work to the CUDA back end
 * through the C ABI in include/mtgl_dev.h.
 */
#ifndef MTGL_B200_GL_H
#define MTGL_B200_GL_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* scalar types (reference: include/GL/gl.h:18-34) */
typedef void           GLvoid;
typedef unsigned char  GLboolean;
typedef signed char    GLbyte;
typedef unsigned char  GLubyte;
typedef short          GLshort;
typedef unsigned short GLushort;
typedef int            GLint;
typedef unsigned int   GLuint;
typedef int            GLsizei;
typedef unsigned int   GLenum;
typedef unsigned int   GLbitfield;
typedef float          GLfloat;
typedef float          GLclampf;
typedef double         GLdouble;
typedef double         GLclampd;
typedef long           GLsizeiptr;
typedef long           GLintptr;

#define GL_FALSE 0
#define GL_TRUE  1

/* ---- glClear / attribute bit masks ---- */
#define GL_DEPTH_BUFFER_BIT              0x00000100
#define GL_STENCIL_BUFFER_BIT            0x00000400
#define GL_COLOR_BUFFER_BIT              0x00004000

/* ---- enumerants, ordered by token value (Khronos registry numbering) ---- */
/* 0x00xx */
#define GL_NO_ERROR                              0x0000
#define GL_POINTS                                0x0000
#define GL_ZERO                                  0x0000
#define GL_LINES                                 0x0001
#define GL_ONE                                   0x0001
#define GL_LINE_LOOP                             0x0002
#define GL_LINE_STRIP                            0x0003
#define GL_TRIANGLES                             0x0004
#define GL_TRIANGLE_STRIP                        0x0005
#define GL_TRIANGLE_FAN                          0x0006
#define GL_QUADS                                 0x0007
#define GL_QUAD_STRIP                            0x0008
#define GL_POLYGON                               0x0009
/* 0x01xx */
#define GL_ADD                                   0x0104
/* 0x02xx */
#define GL_NEVER                                 0x0200
#define GL_LESS                                  0x0201
#define GL_EQUAL                                 0x0202
#define GL_LEQUAL                                0x0203
#define GL_GREATER                               0x0204
#define GL_NOTEQUAL                              0x0205
#define GL_GEQUAL                                0x0206
#define GL_ALWAYS                                0x0207
/* 0x03xx */
#define GL_SRC_COLOR                             0x0300
#define GL_ONE_MINUS_SRC_COLOR                   0x0301
#define GL_SRC_ALPHA                             0x0302
#define GL_ONE_MINUS_SRC_ALPHA                   0x0303
#define GL_DST_ALPHA                             0x0304
#define GL_ONE_MINUS_DST_ALPHA                   0x0305
#define GL_DST_COLOR                             0x0306
#define GL_ONE_MINUS_DST_COLOR                   0x0307
#define GL_SRC_ALPHA_SATURATE                    0x0308
/* 0x04xx */
#define GL_FRONT                                 0x0404
#define GL_BACK                                  0x0405
#define GL_FRONT_AND_BACK                        0x0408
/* 0x05xx */
#define GL_INVALID_ENUM                          0x0500
#define GL_INVALID_VALUE                         0x0501
#define GL_INVALID_OPERATION                     0x0502
#define GL_STACK_OVERFLOW                        0x0503
#define GL_STACK_UNDERFLOW                       0x0504
#define GL_OUT_OF_MEMORY                         0x0505
/* 0x08xx */
#define GL_EXP                                   0x0800
#define GL_EXP2                                  0x0801
/* 0x09xx */
#define GL_CW                                    0x0900
#define GL_CCW                                   0x0901
/* 0x0bxx */
#define GL_CURRENT_COLOR                         0x0b00
#define GL_CURRENT_INDEX                         0x0b01
#define GL_CURRENT_NORMAL                        0x0b02
#define GL_CURRENT_TEXTURE_COORDS                0x0b03
#define GL_CURRENT_RASTER_COLOR                  0x0b04
#define GL_CURRENT_RASTER_INDEX                  0x0b05
#define GL_CURRENT_RASTER_TEXTURE_COORDS         0x0b06
#define GL_CURRENT_RASTER_POSITION               0x0b07
#define GL_CURRENT_RASTER_POSITION_VALID         0x0b08
#define GL_CURRENT_RASTER_DISTANCE               0x0b09
#define GL_POINT_SMOOTH                          0x0b10
#define GL_POINT_SIZE                            0x0b11
#define GL_POINT_SIZE_RANGE                      0x0b12
#define GL_POINT_SIZE_GRANULARITY                0x0b13
#define GL_LINE_SMOOTH                           0x0b20
#define GL_LINE_WIDTH                            0x0b21
#define GL_LINE_WIDTH_RANGE                      0x0b22
#define GL_LINE_WIDTH_GRANULARITY                0x0b23
#define GL_LINE_STIPPLE                          0x0b24
#define GL_LINE_STIPPLE_PATTERN                  0x0b25
#define GL_LINE_STIPPLE_REPEAT                   0x0b26
#define GL_LIST_MODE                             0x0b30
#define GL_LIST_BASE                             0x0b32
#define GL_LIST_INDEX                            0x0b33
#define GL_POLYGON_MODE                          0x0b40
#define GL_POLYGON_SMOOTH                        0x0b41
#define GL_POLYGON_STIPPLE                       0x0b42
#define GL_CULL_FACE                             0x0b44
#define GL_CULL_FACE_MODE                        0x0b45
#define GL_FRONT_FACE                            0x0b46
#define GL_LIGHTING                              0x0b50
#define GL_LIGHT_MODEL_LOCAL_VIEWER              0x0b51
#define GL_LIGHT_MODEL_TWO_SIDE                  0x0b52
#define GL_LIGHT_MODEL_AMBIENT                   0x0b53
#define GL_SHADE_MODEL                           0x0b54
#define GL_COLOR_MATERIAL_FACE                   0x0b55
#define GL_COLOR_MATERIAL_PARAMETER              0x0b56
#define GL_COLOR_MATERIAL                        0x0b57
#define GL_FOG                                   0x0b60
#define GL_FOG_INDEX                             0x0b61
#define GL_FOG_DENSITY                           0x0b62
#define GL_FOG_START                             0x0b63
#define GL_FOG_END                               0x0b64
#define GL_FOG_MODE                              0x0b65
#define GL_FOG_COLOR                             0x0b66
#define GL_DEPTH_RANGE                           0x0b70
#define GL_DEPTH_TEST                            0x0b71
#define GL_DEPTH_WRITEMASK                       0x0b72
#define GL_DEPTH_CLEAR_VALUE                     0x0b73
#define GL_DEPTH_FUNC                            0x0b74
#define GL_STENCIL_TEST                          0x0b90
#define GL_STENCIL_CLEAR_VALUE                   0x0b91
#define GL_STENCIL_FUNC                          0x0b92
#define GL_STENCIL_VALUE_MASK                    0x0b93
#define GL_STENCIL_FAIL                          0x0b94
#define GL_STENCIL_PASS_DEPTH_FAIL               0x0b95
#define GL_STENCIL_PASS_DEPTH_PASS               0x0b96
#define GL_STENCIL_REF                           0x0b97
#define GL_STENCIL_WRITEMASK                     0x0b98
#define GL_MATRIX_MODE                           0x0ba0
#define GL_NORMALIZE                             0x0ba1
#define GL_VIEWPORT                              0x0ba2
#define GL_MODELVIEW_STACK_DEPTH                 0x0ba3
#define GL_PROJECTION_STACK_DEPTH                0x0ba4
#define GL_TEXTURE_STACK_DEPTH                   0x0ba5
#define GL_MODELVIEW_MATRIX                      0x0ba6
#define GL_PROJECTION_MATRIX                     0x0ba7
#define GL_TEXTURE_MATRIX                        0x0ba8
#define GL_ATTRIB_STACK_DEPTH                    0x0bb0
#define GL_CLIENT_ATTRIB_STACK_DEPTH             0x0bb1
#define GL_ALPHA_TEST                            0x0bc0
#define GL_ALPHA_TEST_FUNC                       0x0bc1
#define GL_ALPHA_TEST_REF                        0x0bc2
#define GL_DITHER                                0x0bd0
#define GL_BLEND_DST                             0x0be0
#define GL_BLEND_SRC                             0x0be1
#define GL_BLEND                                 0x0be2
#define GL_LOGIC_OP_MODE                         0x0bf0
#define GL_INDEX_LOGIC_OP                        0x0bf1
#define GL_COLOR_LOGIC_OP                        0x0bf2
/* 0x0cxx */
#define GL_AUX_BUFFERS                           0x0c00
#define GL_DRAW_BUFFER                           0x0c01
#define GL_READ_BUFFER                           0x0c02
#define GL_SCISSOR_BOX                           0x0c10
#define GL_SCISSOR_TEST                          0x0c11
#define GL_INDEX_WRITEMASK                       0x0c21
#define GL_COLOR_CLEAR_VALUE                     0x0c22
#define GL_COLOR_WRITEMASK                       0x0c23
#define GL_INDEX_MODE                            0x0c30
#define GL_RGBA_MODE                             0x0c31
#define GL_DOUBLEBUFFER                          0x0c32
#define GL_STEREO                                0x0c33
#define GL_RENDER_MODE                           0x0c40
#define GL_PERSPECTIVE_CORRECTION_HINT           0x0c50
#define GL_POINT_SMOOTH_HINT                     0x0c51
#define GL_LINE_SMOOTH_HINT                      0x0c52
#define GL_POLYGON_SMOOTH_HINT                   0x0c53
#define GL_FOG_HINT                              0x0c54
#define GL_TEXTURE_GEN_S                         0x0c60
#define GL_TEXTURE_GEN_T                         0x0c61
#define GL_TEXTURE_GEN_R                         0x0c62
#define GL_TEXTURE_GEN_Q                         0x0c63
#define GL_UNPACK_SWAP_BYTES                     0x0cf0
#define GL_UNPACK_LSB_FIRST                      0x0cf1
#define GL_UNPACK_ROW_LENGTH                     0x0cf2
#define GL_UNPACK_SKIP_ROWS                      0x0cf3
#define GL_UNPACK_SKIP_PIXELS                    0x0cf4
#define GL_UNPACK_ALIGNMENT                      0x0cf5
/* 0x0dxx */
#define GL_PACK_SWAP_BYTES                       0x0d00
#define GL_PACK_LSB_FIRST                        0x0d01
#define GL_PACK_ROW_LENGTH                       0x0d02
#define GL_PACK_SKIP_ROWS                        0x0d03
#define GL_PACK_SKIP_PIXELS                      0x0d04
#define GL_PACK_ALIGNMENT                        0x0d05
#define GL_MAP_COLOR                             0x0d10
#define GL_MAP_STENCIL                           0x0d11
#define GL_INDEX_SHIFT                           0x0d12
#define GL_INDEX_OFFSET                          0x0d13
#define GL_RED_SCALE                             0x0d14
#define GL_RED_BIAS                              0x0d15
#define GL_ZOOM_X                                0x0d16
#define GL_ZOOM_Y                                0x0d17
#define GL_GREEN_SCALE                           0x0d18
#define GL_GREEN_BIAS                            0x0d19
#define GL_BLUE_SCALE                            0x0d1a
#define GL_BLUE_BIAS                             0x0d1b
#define GL_ALPHA_SCALE                           0x0d1c
#define GL_ALPHA_BIAS                            0x0d1d
#define GL_DEPTH_SCALE                           0x0d1e
#define GL_DEPTH_BIAS                            0x0d1f
#define GL_MAX_LIGHTS                            0x0d31
#define GL_MAX_CLIP_PLANES                       0x0d32
#define GL_MAX_TEXTURE_SIZE                      0x0d33
#define GL_MAX_PIXEL_MAP_TABLE                   0x0d34
#define GL_MAX_ATTRIB_STACK_DEPTH                0x0d35
#define GL_MAX_MODELVIEW_STACK_DEPTH             0x0d36
#define GL_MAX_NAME_STACK_DEPTH                  0x0d37
#define GL_MAX_PROJECTION_STACK_DEPTH            0x0d38
#define GL_MAX_TEXTURE_STACK_DEPTH               0x0d39
#define GL_MAX_VIEWPORT_DIMS                     0x0d3a
#define GL_MAX_CLIENT_ATTRIB_STACK_DEPTH         0x0d3b
#define GL_SUBPIXEL_BITS                         0x0d50
#define GL_INDEX_BITS                            0x0d51
#define GL_RED_BITS                              0x0d52
#define GL_GREEN_BITS                            0x0d53
#define GL_BLUE_BITS                             0x0d54
#define GL_ALPHA_BITS                            0x0d55
#define GL_DEPTH_BITS                            0x0d56
#define GL_STENCIL_BITS                          0x0d57
#define GL_ACCUM_RED_BITS                        0x0d58
#define GL_ACCUM_GREEN_BITS                      0x0d59
#define GL_ACCUM_BLUE_BITS                       0x0d5a
#define GL_ACCUM_ALPHA_BITS                      0x0d5b
#define GL_NAME_STACK_DEPTH                      0x0d70
#define GL_TEXTURE_1D                            0x0de0
#define GL_TEXTURE_2D                            0x0de1
#define GL_FEEDBACK_BUFFER_SIZE                  0x0df1
#define GL_FEEDBACK_BUFFER_TYPE                  0x0df2
#define GL_SELECTION_BUFFER_SIZE                 0x0df4
/* 0x11xx */
#define GL_DONT_CARE                             0x1100
#define GL_FASTEST                               0x1101
#define GL_NICEST                                0x1102
/* 0x12xx */
#define GL_AMBIENT                               0x1200
#define GL_DIFFUSE                               0x1201
#define GL_SPECULAR                              0x1202
#define GL_POSITION                              0x1203
#define GL_SPOT_DIRECTION                        0x1204
#define GL_SPOT_EXPONENT                         0x1205
#define GL_SPOT_CUTOFF                           0x1206
#define GL_CONSTANT_ATTENUATION                  0x1207
#define GL_LINEAR_ATTENUATION                    0x1208
#define GL_QUADRATIC_ATTENUATION                 0x1209
/* 0x13xx */
#define GL_COMPILE                               0x1300
#define GL_COMPILE_AND_EXECUTE                   0x1301
/* 0x14xx */
#define GL_UNSIGNED_BYTE                         0x1401
#define GL_UNSIGNED_SHORT                        0x1403
#define GL_UNSIGNED_INT                          0x1405
#define GL_FLOAT                                 0x1406
/* 0x15xx */
#define GL_INVERT                                0x150a
/* 0x16xx */
#define GL_EMISSION                              0x1600
#define GL_SHININESS                             0x1601
#define GL_AMBIENT_AND_DIFFUSE                   0x1602
#define GL_COLOR_INDEXES                         0x1603
/* 0x17xx */
#define GL_MODELVIEW                             0x1700
#define GL_PROJECTION                            0x1701
#define GL_TEXTURE                               0x1702
/* 0x19xx */
#define GL_RGB                                   0x1907
#define GL_RGBA                                  0x1908
#define GL_LUMINANCE                             0x1909
#define GL_LUMINANCE_ALPHA                       0x190a
/* 0x1bxx */
#define GL_POINT                                 0x1b00
#define GL_LINE                                  0x1b01
#define GL_FILL                                  0x1b02
/* 0x1dxx */
#define GL_FLAT                                  0x1d00
#define GL_SMOOTH                                0x1d01
#define GL_PHONG                                 0x1d02
/* 0x1exx */
#define GL_KEEP                                  0x1e00
#define GL_REPLACE                               0x1e01
#define GL_INCR                                  0x1e02
#define GL_DECR                                  0x1e03
/* 0x1fxx */
#define GL_VENDOR                                0x1f00
#define GL_RENDERER                              0x1f01
#define GL_VERSION                               0x1f02
#define GL_EXTENSIONS                            0x1f03
/* 0x21xx */
#define GL_MODULATE                              0x2100
#define GL_DECAL                                 0x2101
/* 0x22xx */
#define GL_TEXTURE_ENV_MODE                      0x2200
#define GL_TEXTURE_ENV_COLOR                     0x2201
/* 0x23xx */
#define GL_TEXTURE_ENV                           0x2300
/* 0x26xx */
#define GL_NEAREST                               0x2600
#define GL_LINEAR                                0x2601
/* 0x27xx */
#define GL_NEAREST_MIPMAP_NEAREST                0x2700
#define GL_LINEAR_MIPMAP_NEAREST                 0x2701
#define GL_NEAREST_MIPMAP_LINEAR                 0x2702
#define GL_LINEAR_MIPMAP_LINEAR                  0x2703
/* 0x28xx */
#define GL_TEXTURE_MAG_FILTER                    0x2800
#define GL_TEXTURE_MIN_FILTER                    0x2801
#define GL_TEXTURE_WRAP_S                        0x2802
#define GL_TEXTURE_WRAP_T                        0x2803
/* 0x29xx */
#define GL_CLAMP                                 0x2900
#define GL_REPEAT                                0x2901
/* 0x2axx */
#define GL_POLYGON_OFFSET_UNITS                  0x2a00
#define GL_POLYGON_OFFSET_POINT                  0x2a01
#define GL_POLYGON_OFFSET_LINE                   0x2a02
/* 0x40xx */
#define GL_LIGHT0                                0x4000
#define GL_LIGHT1                                0x4001
#define GL_LIGHT2                                0x4002
#define GL_LIGHT3                                0x4003
#define GL_LIGHT4                                0x4004
#define GL_LIGHT5                                0x4005
#define GL_LIGHT6                                0x4006
#define GL_LIGHT7                                0x4007
/* 0x80xx */
#define GL_CONSTANT_COLOR                        0x8001
#define GL_ONE_MINUS_CONSTANT_COLOR              0x8002
#define GL_CONSTANT_ALPHA                        0x8003
#define GL_ONE_MINUS_CONSTANT_ALPHA              0x8004
#define GL_BLEND_COLOR                           0x8005
#define GL_POLYGON_OFFSET_FILL                   0x8037
#define GL_POLYGON_OFFSET_FACTOR                 0x8038
#define GL_RESCALE_NORMAL                        0x803a
#define GL_TEXTURE_BINDING_1D                    0x8068
#define GL_TEXTURE_BINDING_2D                    0x8069
#define GL_TEXTURE_BINDING_3D                    0x806a
#define GL_TEXTURE_3D                            0x806f
#define GL_MAX_3D_TEXTURE_SIZE                   0x8073
#define GL_VERTEX_ARRAY                          0x8074
#define GL_NORMAL_ARRAY                          0x8075
#define GL_COLOR_ARRAY                           0x8076
#define GL_INDEX_ARRAY                           0x8077
#define GL_TEXTURE_COORD_ARRAY                   0x8078
#define GL_EDGE_FLAG_ARRAY                       0x8079
#define GL_VERTEX_ARRAY_SIZE                     0x807a
#define GL_VERTEX_ARRAY_TYPE                     0x807b
#define GL_VERTEX_ARRAY_STRIDE                   0x807c
#define GL_NORMAL_ARRAY_TYPE                     0x807e
#define GL_NORMAL_ARRAY_STRIDE                   0x807f
#define GL_COLOR_ARRAY_SIZE                      0x8081
#define GL_COLOR_ARRAY_TYPE                      0x8082
#define GL_COLOR_ARRAY_STRIDE                    0x8083
#define GL_TEXTURE_COORD_ARRAY_SIZE              0x8088
#define GL_TEXTURE_COORD_ARRAY_TYPE              0x8089
#define GL_TEXTURE_COORD_ARRAY_STRIDE            0x808a
#define GL_SAMPLE_BUFFERS                        0x80a8
#define GL_SAMPLES                               0x80a9
#define GL_SAMPLE_COVERAGE_VALUE                 0x80aa
#define GL_SAMPLE_COVERAGE_INVERT                0x80ab
#define GL_BLEND_DST_RGB                         0x80c8
#define GL_BLEND_SRC_RGB                         0x80c9
#define GL_BLEND_DST_ALPHA                       0x80ca
#define GL_BLEND_SRC_ALPHA                       0x80cb
#define GL_MAX_ELEMENTS_VERTICES                 0x80e8
#define GL_MAX_ELEMENTS_INDICES                  0x80e9
/* 0x81xx */
#define GL_POINT_SIZE_MIN                        0x8126
#define GL_POINT_SIZE_MAX                        0x8127
#define GL_POINT_FADE_THRESHOLD_SIZE             0x8128
#define GL_POINT_DISTANCE_ATTENUATION            0x8129
#define GL_CLAMP_TO_EDGE                         0x812f
#define GL_GENERATE_MIPMAP_HINT                  0x8192
#define GL_LIGHT_MODEL_COLOR_CONTROL             0x81f8
#define GL_SINGLE_COLOR                          0x81f9
#define GL_SEPARATE_SPECULAR_COLOR               0x81fa
/* 0x84xx */
#define GL_MAX_TEXTURE_UNITS                     0x84e2
#define GL_MAX_TEXTURE_LOD_BIAS                  0x84fd
/* 0x85xx */
#define GL_INCR_WRAP                             0x8507
#define GL_DECR_WRAP                             0x8508
#define GL_MAX_CUBE_MAP_TEXTURE_SIZE             0x851c
/* 0x87xx */
#define GL_BUFFER_SIZE                           0x8764
#define GL_BUFFER_USAGE                          0x8765
/* 0x88xx */
#define GL_ARRAY_BUFFER                          0x8892
#define GL_ELEMENT_ARRAY_BUFFER                  0x8893
#define GL_ARRAY_BUFFER_BINDING                  0x8894
#define GL_ELEMENT_ARRAY_BUFFER_BINDING          0x8895
#define GL_VERTEX_ARRAY_BUFFER_BINDING           0x8896
#define GL_NORMAL_ARRAY_BUFFER_BINDING           0x8897
#define GL_COLOR_ARRAY_BUFFER_BINDING            0x8898
#define GL_TEXTURE_COORD_ARRAY_BUFFER_BINDING    0x889a
#define GL_BUFFER_ACCESS                         0x88bb
#define GL_BUFFER_MAPPED                         0x88bc
#define GL_BUFFER_MAP_POINTER                    0x88bd
#define GL_STREAM_DRAW                           0x88e0
#define GL_STREAM_READ                           0x88e1
#define GL_STREAM_COPY                           0x88e2
#define GL_STATIC_DRAW                           0x88e4
#define GL_STATIC_READ                           0x88e5
#define GL_STATIC_COPY                           0x88e6
#define GL_DYNAMIC_DRAW                          0x88e8
#define GL_DYNAMIC_READ                          0x88e9
#define GL_DYNAMIC_COPY                          0x88ea

/* ------------------------------------------------------------------------------------------
 * Entry points.  Grouped by the pipeline stage whose state they feed.
 * ------------------------------------------------------------------------------------------ */

/* capabilities, clears, viewport */
void glEnable(GLenum cap);
void glDisable(GLenum cap);
GLboolean glIsEnabled(GLenum cap);
void glClear(GLbitfield mask);
void glClearColor(GLclampf r, GLclampf g, GLclampf b, GLclampf a);
void glClearDepth(GLclampd depth);
void glClearStencil(GLint s);
void glViewport(GLint x, GLint y, GLsizei w, GLsizei h);
void glScissor(GLint x, GLint y, GLsizei w, GLsizei h);
void glDepthRange(GLclampd zNear, GLclampd zFar);

/* matrix stacks */
void glMatrixMode(GLenum mode);
void glLoadIdentity(void);
void glLoadMatrixf(const GLfloat *m);
void glMultMatrixf(const GLfloat *m);
void glPushMatrix(void);
void glPopMatrix(void);
void glTranslatef(GLfloat x, GLfloat y, GLfloat z);
void glRotatef(GLfloat angle, GLfloat x, GLfloat y, GLfloat z);
void glScalef(GLfloat x, GLfloat y, GLfloat z);
void glOrtho(GLdouble l, GLdouble r, GLdouble b, GLdouble t, GLdouble zNear, GLdouble zFar);
void glFrustum(GLdouble l, GLdouble r, GLdouble b, GLdouble t, GLdouble zNear, GLdouble zFar);

/* immediate mode */
void glBegin(GLenum mode);
void glEnd(void);
void glVertex2f(GLfloat x, GLfloat y);
void glVertex3f(GLfloat x, GLfloat y, GLfloat z);
void glVertex2i(GLint x, GLint y);
void glVertex3i(GLint x, GLint y, GLint z);
void glColor3f(GLfloat r, GLfloat g, GLfloat b);
void glColor4f(GLfloat r, GLfloat g, GLfloat b, GLfloat a);
void glColor3ub(GLubyte r, GLubyte g, GLubyte b);
void glColor4ub(GLubyte r, GLubyte g, GLubyte b, GLubyte a);
void glTexCoord2f(GLfloat s, GLfloat t);
void glNormal3f(GLfloat nx, GLfloat ny, GLfloat nz);

/* vertex arrays and buffer objects (GL 1.1 / 1.5) */
void glEnableClientState(GLenum array);
void glDisableClientState(GLenum array);
void glVertexPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *ptr);
void glColorPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *ptr);
void glTexCoordPointer(GLint size, GLenum type, GLsizei stride, const GLvoid *ptr);
void glNormalPointer(GLenum type, GLsizei stride, const GLvoid *ptr);
void glDrawArrays(GLenum mode, GLint first, GLsizei count);
void glDrawElements(GLenum mode, GLsizei count, GLenum type, const GLvoid *indices);
void glGenBuffers(GLsizei n, GLuint *ids);
void glDeleteBuffers(GLsizei n, const GLuint *ids);
void glBindBuffer(GLenum target, GLuint id);
void glBufferData(GLenum target, GLsizeiptr size, const GLvoid *data, GLenum usage);
void glBufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const GLvoid *data);
GLboolean glIsBuffer(GLuint id);

/* textures and texture environment */
void glGenTextures(GLsizei n, GLuint *ids);
void glDeleteTextures(GLsizei n, const GLuint *ids);
void glBindTexture(GLenum target, GLuint id);
void glTexImage2D(GLenum target, GLint level, GLint internalformat, GLsizei w, GLsizei h,
                  GLint border, GLenum format, GLenum type, const GLvoid *pixels);
void glTexParameteri(GLenum target, GLenum pname, GLint param);
void glTexEnvi(GLenum target, GLenum pname, GLint param);
void glTexEnvf(GLenum target, GLenum pname, GLfloat param);
void glTexEnvfv(GLenum target, GLenum pname, const GLfloat *params);
GLboolean glIsTexture(GLuint id);

/* lighting and materials */
void glShadeModel(GLenum mode);
void glLightf(GLenum light, GLenum pname, GLfloat param);
void glLighti(GLenum light, GLenum pname, GLint param);
void glLightfv(GLenum light, GLenum pname, const GLfloat *params);
void glLightiv(GLenum light, GLenum pname, const GLint *params);
void glLightModelf(GLenum pname, GLfloat param);
void glLightModeli(GLenum pname, GLint param);
void glLightModelfv(GLenum pname, const GLfloat *params);
void glLightModeliv(GLenum pname, const GLint *params);
void glMaterialf(GLenum face, GLenum pname, GLfloat param);
void glMateriali(GLenum face, GLenum pname, GLint param);
void glMaterialfv(GLenum face, GLenum pname, const GLfloat *params);
void glMaterialiv(GLenum face, GLenum pname, const GLint *params);
void glColorMaterial(GLenum face, GLenum mode);
void glGetLightfv(GLenum light, GLenum pname, GLfloat *params);
void glGetMaterialfv(GLenum face, GLenum pname, GLfloat *params);

/* fog */
void glFogi(GLenum pname, GLint param);
void glFogf(GLenum pname, GLfloat param);
void glFogfv(GLenum pname, const GLfloat *params);

/* rasterisation and per-fragment state */
void glCullFace(GLenum mode);
void glFrontFace(GLenum mode);
void glPolygonMode(GLenum face, GLenum mode);
void glLineWidth(GLfloat width);
void glPointSize(GLfloat size);
void glHint(GLenum target, GLenum mode);
void glAlphaFunc(GLenum func, GLclampf ref);
void glStencilFunc(GLenum func, GLint ref, GLuint mask);
void glStencilOp(GLenum sfail, GLenum dpfail, GLenum dppass);
void glStencilMask(GLuint mask);
void glDepthFunc(GLenum func);
void glDepthMask(GLboolean flag);
void glBlendFunc(GLenum sfactor, GLenum dfactor);
void glColorMask(GLboolean r, GLboolean g, GLboolean b, GLboolean a);

/* pixel rectangles */
void glPixelStorei(GLenum pname, GLint param);
void glReadPixels(GLint x, GLint y, GLsizei w, GLsizei h, GLenum format, GLenum type, GLvoid *pixels);
void glDrawPixels(GLsizei w, GLsizei h, GLenum format, GLenum type, const GLvoid *pixels);
void glRasterPos2i(GLint x, GLint y);
void glRasterPos2f(GLfloat x, GLfloat y);
void glRasterPos3f(GLfloat x, GLfloat y, GLfloat z);

/* display lists */
GLuint glGenLists(GLsizei range);
void glDeleteLists(GLuint list, GLsizei range);
void glNewList(GLuint list, GLenum mode);
void glEndList(void);
void glCallList(GLuint list);
void glCallLists(GLsizei n, GLenum type, const GLvoid *lists);
void glListBase(GLuint base);
GLboolean glIsList(GLuint list);

/* queries and synchronisation */
void glGetIntegerv(GLenum pname, GLint *params);
void glGetFloatv(GLenum pname, GLfloat *params);
void glGetDoublev(GLenum pname, GLdouble *params);
void glGetBooleanv(GLenum pname, GLboolean *params);
const GLubyte *glGetString(GLenum name);
GLenum glGetError(void);
void glFlush(void);
void glFinish(void);

#ifdef __cplusplus
}
#endif

#endif /* MTGL_B200_GL_H */
