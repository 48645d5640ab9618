/*
 * scenes.c -- headless scene recipes shared by the parity tests and the benchmark.
 *
 * Uses ONLY the public gl* API (GL/gl.h), so the very same translation unit is compiled twice:
 * into oracle/_ref/libref_*.so together with the unmodified reference sources, and into the
 * product library.  The recipes restate the reference's interactive testbed programs without
 * SDL (testbed/1.0-7, 1.0-9, 1.0-14, 1.0-15, 1.0-17, 1.0-18, 1.5-0 ...) plus the five BASELINE.json
 * configurations as fixed by SURVEY.md section 8(d).
 *
 * The Suzanne mesh is not stored here: the harness passes it in (tests/golden/suzanne.npz,
 * extracted from the reference's testbed/suzanne_data.h by tests/golden/make_fixtures.py).
 */
#include <GL/gl.h>

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- mesh supplied by the harness */
static float *g_mesh_pos;   /* nv * 3 */
static float *g_mesh_nrm;   /* nv * 3 */
static int *g_mesh_faces;   /* nf * 3 vertex indices (normal index == vertex index) */
static int g_mesh_nv, g_mesh_nf;

void scene_set_mesh(const float *pos, const float *nrm, const int *faces, int nv, int nf)
{
    free(g_mesh_pos); free(g_mesh_nrm); free(g_mesh_faces);
    g_mesh_pos = (float *)malloc(sizeof(float) * 3 * (size_t)nv);
    g_mesh_nrm = (float *)malloc(sizeof(float) * 3 * (size_t)nv);
    g_mesh_faces = (int *)malloc(sizeof(int) * 3 * (size_t)nf);
    memcpy(g_mesh_pos, pos, sizeof(float) * 3 * (size_t)nv);
    memcpy(g_mesh_nrm, nrm, sizeof(float) * 3 * (size_t)nv);
    memcpy(g_mesh_faces, faces, sizeof(int) * 3 * (size_t)nf);
    g_mesh_nv = nv; g_mesh_nf = nf;
}

/* ---------------------------------------------------------------- procedural textures */
static GLuint make_checker_rgb(int size, int cell, int white_first) /* 1.0-7:25-42 / 1.0-14:50-64 */
{
    uint8_t *px = (uint8_t *)malloc((size_t)size * size * 3);
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            int on = ((x / cell) + (y / cell)) % 2;
            if (white_first) on = !on;
            uint8_t *p = px + ((size_t)y * size + x) * 3;
            if (on) { p[0] = 255; p[1] = 255; p[2] = 255; } else { p[0] = 50; p[1] = 50; p[2] = 200; }
        }
    GLuint id;
    glGenTextures(1, &id);
    glBindTexture(GL_TEXTURE_2D, id);
    glTexImage2D(GL_TEXTURE_2D, 0, GL_RGB, size, size, 0, GL_RGB, GL_UNSIGNED_BYTE, px);
    free(px);
    return id;
}

static GLuint make_radial_rgba(int size) /* 1.0-15:40-73 */
{
    uint8_t *px = (uint8_t *)malloc((size_t)size * size * 4);
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            uint8_t *p = px + ((size_t)y * size + x) * 4;
            float cx = (x - size / 2.0f) / (size / 2.0f);
            float cy = (y - size / 2.0f) / (size / 2.0f);
            float dist = sqrtf(cx * cx + cy * cy);
            if (dist < 0.8f) {
                float alpha = 1.0f - (dist / 0.8f);
                p[0] = 255; p[1] = 255; p[2] = 0; p[3] = (uint8_t)(alpha * 255);
            } else { p[0] = p[1] = p[2] = p[3] = 0; }
        }
    GLuint id;
    glGenTextures(1, &id);
    glBindTexture(GL_TEXTURE_2D, id);
    glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA, size, size, 0, GL_RGBA, GL_UNSIGNED_BYTE, px);
    free(px);
    return id;
}

static GLuint make_gradient_la(int w, int h) /* luminance+alpha ramp: exercises the LA upload and non-square sizes */
{
    uint8_t *px = (uint8_t *)malloc((size_t)w * h * 2);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            px[((size_t)y * w + x) * 2] = (uint8_t)((x * 255) / (w > 1 ? w - 1 : 1));
            px[((size_t)y * w + x) * 2 + 1] = (uint8_t)(255 - (y * 200) / (h > 1 ? h - 1 : 1));
        }
    GLuint id;
    glGenTextures(1, &id);
    glBindTexture(GL_TEXTURE_2D, id);
    glTexImage2D(GL_TEXTURE_2D, 0, GL_LUMINANCE_ALPHA, w, h, 0, GL_LUMINANCE_ALPHA, GL_UNSIGNED_BYTE, px);
    free(px);
    return id;
}

/* ---------------------------------------------------------------- helpers */
static void frustum_like_testbed(int w, int h, double zfar) /* 1.0-18:95-101 */
{
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    float aspect = (float)w / (float)h;
    glFrustum(-aspect * 0.1, aspect * 0.1, -0.1, 0.1, 0.1, zfar);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
}

static void suzanne_immediate(void) /* 1.0-18:25-41 */
{
    glBegin(GL_TRIANGLES);
    for (int i = 0; i < g_mesh_nf; i++)
        for (int j = 0; j < 3; j++) {
            int vi = g_mesh_faces[i * 3 + j];
            glNormal3f(g_mesh_nrm[vi * 3], g_mesh_nrm[vi * 3 + 1], g_mesh_nrm[vi * 3 + 2]);
            glVertex3f(g_mesh_pos[vi * 3], g_mesh_pos[vi * 3 + 1], g_mesh_pos[vi * 3 + 2]);
        }
    glEnd();
}

static void suzanne_lights_and_material(void) /* 1.0-18:107-138 */
{
    glEnable(GL_LIGHTING);
    glEnable(GL_LIGHT0);
    glEnable(GL_LIGHT1);
    GLfloat l0p[] = { 3.0f, 3.0f, 3.0f, 0.0f }, l0d[] = { 1.0f, 0.95f, 0.9f, 1.0f }, l0s[] = { 1, 1, 1, 1 };
    glLightfv(GL_LIGHT0, GL_POSITION, l0p);
    glLightfv(GL_LIGHT0, GL_DIFFUSE, l0d);
    glLightfv(GL_LIGHT0, GL_SPECULAR, l0s);
    GLfloat l1p[] = { -2.0f, -1.0f, 2.0f, 0.0f }, l1d[] = { 0.3f, 0.4f, 0.5f, 1.0f };
    glLightfv(GL_LIGHT1, GL_POSITION, l1p);
    glLightfv(GL_LIGHT1, GL_DIFFUSE, l1d);
    GLfloat amb[] = { 0.15f, 0.15f, 0.2f, 1.0f };
    glLightModelfv(GL_LIGHT_MODEL_AMBIENT, amb);
    GLfloat ma[] = { 0.3f, 0.2f, 0.1f, 1.0f }, md[] = { 0.8f, 0.5f, 0.3f, 1.0f }, ms[] = { 0.4f, 0.4f, 0.4f, 1.0f };
    glMaterialfv(GL_FRONT_AND_BACK, GL_AMBIENT, ma);
    glMaterialfv(GL_FRONT_AND_BACK, GL_DIFFUSE, md);
    glMaterialfv(GL_FRONT_AND_BACK, GL_SPECULAR, ms);
    glMaterialf(GL_FRONT_AND_BACK, GL_SHININESS, 32.0f);
}

static void cube_quads(float size, float uvs) /* geometry of 1.0-7:45-102, texture coordinates scaled by uvs */
{
    float s = size / 2.0f;
    static const float f[6][4][5] = {
        { { -1, -1, 1, 0, 0 }, { 1, -1, 1, 1, 0 }, { 1, 1, 1, 1, 1 }, { -1, 1, 1, 0, 1 } },
        { { -1, -1, -1, 1, 0 }, { -1, 1, -1, 1, 1 }, { 1, 1, -1, 0, 1 }, { 1, -1, -1, 0, 0 } },
        { { -1, 1, -1, 0, 1 }, { -1, 1, 1, 0, 0 }, { 1, 1, 1, 1, 0 }, { 1, 1, -1, 1, 1 } },
        { { -1, -1, -1, 1, 1 }, { 1, -1, -1, 0, 1 }, { 1, -1, 1, 0, 0 }, { -1, -1, 1, 1, 0 } },
        { { 1, -1, -1, 1, 0 }, { 1, 1, -1, 1, 1 }, { 1, 1, 1, 0, 1 }, { 1, -1, 1, 0, 0 } },
        { { -1, -1, -1, 0, 0 }, { -1, -1, 1, 1, 0 }, { -1, 1, 1, 1, 1 }, { -1, 1, -1, 0, 1 } },
    };
    for (int face = 0; face < 6; face++) {
        glBegin(GL_QUADS);
        glColor3f(1.0f, 1.0f, 1.0f);
        for (int k = 0; k < 4; k++) {
            glTexCoord2f(f[face][k][3] * uvs, f[face][k][4] * uvs);
            glVertex3f(f[face][k][0] * s, f[face][k][1] * s, f[face][k][2] * s);
        }
        glEnd();
    }
}

/* ---------------------------------------------------------------- BASELINE configurations */

/* C1: Suzanne, immediate mode, depth test + lighting.  variant: 0 smooth, 1 phong, 2 flat,
 * 3 smooth + two-sided lighting with a distinct back material, 4 smooth + local viewer */
static void scene_c1(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glDepthFunc(GL_LESS);
    glClearColor(0.2f, 0.2f, 0.3f, 1.0f);
    suzanne_lights_and_material();
    glShadeModel(variant == 1 ? GL_PHONG : variant == 2 ? GL_FLAT : GL_SMOOTH);
    if (variant == 3) {
        GLfloat bd[] = { 0.1f, 0.6f, 0.9f, 1.0f };
        glMaterialfv(GL_BACK, GL_DIFFUSE, bd);
        glLightModeli(GL_LIGHT_MODEL_TWO_SIDE, 1);
    }
    if (variant == 4) glLightModeli(GL_LIGHT_MODEL_LOCAL_VIEWER, 1);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -2.2f);
    glRotatef(20.0f, 1.0f, 0.0f, 0.0f);
    glRotatef(30.0f, 0.0f, 1.0f, 0.0f);
    suzanne_immediate();
}

/* C2: textured cube, trilinear + GL_MODULATE + linear fog, back-face culling */
static void scene_c2_cube(int w, int h, int variant)
{
    (void)variant;
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glEnable(GL_CULL_FACE);
    glCullFace(GL_BACK);
    glClearColor(0.5f, 0.5f, 0.5f, 1.0f);
    glEnable(GL_TEXTURE_2D);
    make_checker_rgb(64, 4, 1);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_REPEAT);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_REPEAT);
    glTexEnvi(GL_TEXTURE_ENV, GL_TEXTURE_ENV_MODE, GL_MODULATE);
    GLfloat fog[] = { 0.5f, 0.5f, 0.5f, 1.0f };
    glEnable(GL_FOG);
    glFogi(GL_FOG_MODE, GL_LINEAR);
    glFogf(GL_FOG_START, 2.0f);
    glFogf(GL_FOG_END, 8.0f);
    glFogfv(GL_FOG_COLOR, fog);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -3.0f);
    glRotatef(30.0f, 1.0f, 0.0f, 0.0f);
    glRotatef(40.0f, 0.0f, 1.0f, 0.0f);
    cube_quads(1.5f, 4.0f);
}

/* C2 sub-scene: the 1.0-14 floor (212-228), real minification and heavy frustum clipping.
 * variant 0..5 = the six minification filters */
static void scene_c2_floor(int w, int h, int variant)
{
    static const GLint filt[6] = { GL_NEAREST, GL_LINEAR, GL_NEAREST_MIPMAP_NEAREST, GL_LINEAR_MIPMAP_NEAREST,
                                   GL_NEAREST_MIPMAP_LINEAR, GL_LINEAR_MIPMAP_LINEAR };
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.1f, 0.1f, 0.2f, 1.0f);
    glEnable(GL_TEXTURE_2D);
    make_checker_rgb(64, 4, 1);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_REPEAT);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_REPEAT);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, filt[variant % 6]);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glLoadIdentity();
    glRotatef(60.0f, 1.0f, 0.0f, 0.0f);
    glTranslatef(0.0f, -2.0f, 0.0f);
    glColor3f(1.0f, 1.0f, 1.0f);
    glBegin(GL_QUADS);
    glTexCoord2f(0.0f, 0.0f);   glVertex3f(-20.0f, 0.0f, -20.0f);
    glTexCoord2f(20.0f, 0.0f);  glVertex3f(20.0f, 0.0f, -20.0f);
    glTexCoord2f(20.0f, 20.0f); glVertex3f(20.0f, 0.0f, 20.0f);
    glTexCoord2f(0.0f, 20.0f);  glVertex3f(-20.0f, 0.0f, 20.0f);
    glEnd();
}

/* C2 sub-scene: the 1.0-15 quad (186-209), one texenv mode per variant (0..4), rotated 17 degrees */
static void scene_c2_texenv(int w, int h, int variant)
{
    static const GLint modes[5] = { GL_MODULATE, GL_DECAL, GL_REPLACE, GL_BLEND, GL_ADD };
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(-2.0, 2.0, -1.5, 1.5, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    glClearColor(0.2f, 0.2f, 0.3f, 1.0f);
    glEnable(GL_TEXTURE_2D);
    make_radial_rgba(64);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_CLAMP);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_CLAMP);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    glTexEnvi(GL_TEXTURE_ENV, GL_TEXTURE_ENV_MODE, modes[variant % 5]);
    GLfloat env[] = { 0.0f, 1.0f, 1.0f, 1.0f };
    glTexEnvfv(GL_TEXTURE_ENV, GL_TEXTURE_ENV_COLOR, env);
    glClear(GL_COLOR_BUFFER_BIT);
    glLoadIdentity();
    glRotatef(17.0f, 0.0f, 0.0f, 1.0f);
    glBegin(GL_QUADS);
    glColor3f(1.0f, 0.0f, 0.0f); glTexCoord2f(0.0f, 0.0f); glVertex2f(-1.0f, -1.0f);
    glColor3f(0.0f, 1.0f, 0.0f); glTexCoord2f(1.0f, 0.0f); glVertex2f(1.0f, -1.0f);
    glColor3f(0.0f, 0.0f, 1.0f); glTexCoord2f(1.0f, 1.0f); glVertex2f(1.0f, 1.0f);
    glColor3f(1.0f, 1.0f, 0.0f); glTexCoord2f(0.0f, 1.0f); glVertex2f(-1.0f, 1.0f);
    glEnd();
}

/* C3: fill-rate stress.  variant bits 0-7 = number of full-screen quads (0 -> 64); bit 8: every quad gets its own texture
 * coordinates (shifted by q / 512), so that no two quads sample the same texels -- the layers are then no longer
 * coincident for the texture stage (a diagnostic variant, not a BASELINE configuration). */
static struct { int nq, shifted; } c3;

/* the state and the texture are set up once per context (scene_c3_setup), scene_c3_draw issues the frame: the clear and
 * the quads -- what bench.py times as a step, like scene_c4_setup / scene_c4_draw */
void scene_c3_setup(int w, int h, int variant)
{
    c3.nq = (variant & 0xFF) > 0 ? (variant & 0xFF) : 64;
    c3.shifted = (variant >> 8) & 1;
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
    glClearStencil(0);
    glEnable(GL_TEXTURE_2D);
    make_radial_rgba(64);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_CLAMP);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_CLAMP);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    glEnable(GL_ALPHA_TEST);
    glAlphaFunc(GL_GREATER, 0.1f);
    glEnable(GL_STENCIL_TEST);
    glStencilFunc(GL_ALWAYS, 0, 0xFF);
    glStencilOp(GL_KEEP, GL_KEEP, GL_INCR_WRAP);
    glEnable(GL_BLEND);
    glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA);
}

void scene_c3_draw(void)
{
    glClear(GL_COLOR_BUFFER_BIT | GL_STENCIL_BUFFER_BIT);
    glBegin(GL_QUADS);
    for (int q = 0; q < c3.nq; q++) {
        glColor4f((float)((q * 37) % 64) / 63.0f, (float)((q * 11) % 64) / 63.0f, (float)((q * 5) % 64) / 63.0f,
                  (q & 1) ? 0.75f : 0.25f);
        float o = c3.shifted ? (float)q / 512.0f : 0.0f;
        glTexCoord2f(0.0f + o, 0.0f); glVertex2f(-1.0f, -1.0f);
        glTexCoord2f(1.0f + o, 0.0f); glVertex2f(1.0f, -1.0f);
        glTexCoord2f(1.0f + o, 1.0f); glVertex2f(1.0f, 1.0f);
        glTexCoord2f(0.0f + o, 1.0f); glVertex2f(-1.0f, 1.0f);
    }
    glEnd();
}

static void scene_c3(int w, int h, int variant)
{
    scene_c3_setup(w, h, variant);
    scene_c3_draw();
}

/* C4 / C5: grid of Suzannes through one interleaved VBO (pos3 normal3 uv2 = 32 B / vertex).
 * The VBO and texture are built once per context by scene_c4_setup; scene_c4_draw issues the frame. */
static struct {
    GLuint vbo, tex;
    int gx, gy, nverts;
    int instanced;  /* variant bit 17: one Suzanne in the VBO, drawn gx * gy times under glPushMatrix / glTranslatef */
    float cz;
    float *host;    /* retained so the end-to-end benchmark can re-upload it every frame */
} g_c4;

static void c4_build(int gx, int gy)
{
    size_t nv = (size_t)gx * gy * g_mesh_nf * 3;
    free(g_c4.host);
    g_c4.host = (float *)malloc(nv * 8 * sizeof(float));
    float *o = g_c4.host;
    for (int iy = 0; iy < gy; iy++)
        for (int ix = 0; ix < gx; ix++) {
            float ox = ((float)ix - (float)(gx - 1) * 0.5f) * 2.1f;
            float oy = ((float)iy - (float)(gy - 1) * 0.5f) * 1.5f;
            for (int f = 0; f < g_mesh_nf; f++)
                for (int j = 0; j < 3; j++) {
                    int vi = g_mesh_faces[f * 3 + j];
                    float x = g_mesh_pos[vi * 3], y = g_mesh_pos[vi * 3 + 1], z = g_mesh_pos[vi * 3 + 2];
                    o[0] = x + ox; o[1] = y + oy; o[2] = z;
                    o[3] = g_mesh_nrm[vi * 3]; o[4] = g_mesh_nrm[vi * 3 + 1]; o[5] = g_mesh_nrm[vi * 3 + 2];
                    o[6] = (x + 0.7f) * 4.0f; o[7] = (y + 0.5f) * 4.0f;
                    o += 8;
                }
        }
    g_c4.gx = gx; g_c4.gy = gy; g_c4.nverts = (int)nv;
}

void scene_c4_upload(void) /* the per-frame host->device part of the end-to-end measurement */
{
    glBindBuffer(GL_ARRAY_BUFFER, g_c4.vbo);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)((size_t)g_c4.nverts * 32), g_c4.host, GL_STATIC_DRAW);
}

/* variant: bits 0-7 grid columns (0 -> 38), bits 8-15 grid rows (0 -> 28), bit 16 GL_PHONG, bit 17 the secondary layout of
 * ONE Suzanne in the VBO (2904 vertices), drawn gx * gy times under glPushMatrix / glTranslatef -- the
 * same picture up to the rounding of "translate, then transform" against "transform the translated vertex" */
void scene_c4_setup(int w, int h, int variant)
{
    int gx = variant & 0xFF, gy = (variant >> 8) & 0xFF;
    if (gx == 0) gx = 38;
    if (gy == 0) gy = 28;
    g_c4.instanced = (variant >> 17) & 1;
    if (g_c4.instanced) { c4_build(1, 1); g_c4.gx = gx; g_c4.gy = gy; }
    else c4_build(gx, gy);
    glGenBuffers(1, &g_c4.vbo);
    scene_c4_upload();
    g_c4.tex = make_checker_rgb(64, 4, 1);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_REPEAT);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_REPEAT);

    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glDepthFunc(GL_LESS);
    glEnable(GL_TEXTURE_2D);
    glClearColor(0.2f, 0.2f, 0.3f, 1.0f);
    glEnable(GL_LIGHTING);
    GLfloat amb[] = { 0.15f, 0.15f, 0.2f, 1.0f };
    glLightModelfv(GL_LIGHT_MODEL_AMBIENT, amb);
    for (int l = 0; l < 8; l++) {   /* even = directional, odd = positional with linear attenuation 0.05 */
        GLfloat pos[4] = { (float)((37 * l) % 7) - 3.0f, (float)((53 * l) % 5) - 2.0f, 3.0f, (l & 1) ? 1.0f : 0.0f };
        GLfloat dif[4] = { 0.10f + 0.05f * (float)(l % 3), 0.12f + 0.04f * (float)((l + 1) % 3), 0.10f + 0.06f * (float)((l + 2) % 3), 1.0f };
        GLfloat spc[4] = { 0.4f, 0.4f, 0.4f, 1.0f };
        glEnable(GL_LIGHT0 + l);
        glLightfv(GL_LIGHT0 + l, GL_POSITION, pos);
        glLightfv(GL_LIGHT0 + l, GL_DIFFUSE, dif);
        glLightfv(GL_LIGHT0 + l, GL_SPECULAR, spc);
        if (l & 1) glLightf(GL_LIGHT0 + l, GL_LINEAR_ATTENUATION, 0.05f);
    }
    GLfloat ma[] = { 0.3f, 0.2f, 0.1f, 1.0f }, md[] = { 0.8f, 0.5f, 0.3f, 1.0f }, ms[] = { 0.4f, 0.4f, 0.4f, 1.0f };
    glMaterialfv(GL_FRONT_AND_BACK, GL_AMBIENT, ma);
    glMaterialfv(GL_FRONT_AND_BACK, GL_DIFFUSE, md);
    glMaterialfv(GL_FRONT_AND_BACK, GL_SPECULAR, ms);
    glMaterialf(GL_FRONT_AND_BACK, GL_SHININESS, 32.0f);
    glShadeModel((variant & (1 << 16)) ? GL_PHONG : GL_SMOOTH);

    glBindBuffer(GL_ARRAY_BUFFER, g_c4.vbo);
    glEnableClientState(GL_VERTEX_ARRAY);
    glEnableClientState(GL_NORMAL_ARRAY);
    glEnableClientState(GL_TEXTURE_COORD_ARRAY);
    glVertexPointer(3, GL_FLOAT, 32, (const void *)0);
    glNormalPointer(GL_FLOAT, 32, (const void *)12);
    glTexCoordPointer(2, GL_FLOAT, 32, (const void *)24);
    glLoadIdentity();
    /* the camera distance of the official layout is 24 for the 38x28 grid; smaller grids move closer
     * so that the per-triangle footprint stays comparable */
    float cz = 24.0f * (float)(gx > gy * 38 / 28 ? gx : gy * 38 / 28) / 38.0f;
    if (cz < 2.5f) cz = 2.5f;
    g_c4.cz = cz;
    glTranslatef(0.0f, 0.0f, -cz);
}

void scene_c4_draw(void)
{
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    if (!g_c4.instanced) { glDrawArrays(GL_TRIANGLES, 0, g_c4.nverts); return; }
    for (int iy = 0; iy < g_c4.gy; iy++)
        for (int ix = 0; ix < g_c4.gx; ix++) {
            glPushMatrix();
            glTranslatef(((float)ix - (float)(g_c4.gx - 1) * 0.5f) * 2.1f, ((float)iy - (float)(g_c4.gy - 1) * 0.5f) * 1.5f, 0.0f);
            glDrawArrays(GL_TRIANGLES, 0, g_c4.nverts);
            glPopMatrix();
        }
}

static void scene_c4(int w, int h, int variant)
{
    scene_c4_setup(w, h, variant);
    scene_c4_draw();
}

/* Geometry far larger than the view, drawn from a static VBO twice: exercises the hierarchical chunk culling in front of
 * set-up (k_cull.cu) -- the second frame reuses the cached chunk boxes on a single device, a band-limited device
 * culls from the first.  The result must not depend on it.
 * variant 0: camera inside a 8x6 grid, most Suzannes off screen on all four sides
 *         1: + rotated about Y so that part of the grid is behind the eye (w <= 0: chunks must not be dropped blindly)
 *         2: + viewport smaller than and offset inside the framebuffer, scissor on
 *         3: GL_LINE polygon mode with wide lines (a triangle then touches pixels outside its vertex box: no culling)
 *         4: position-only array (size 2: z = 0), unlit, flat colour
 *         5: camera close to the grid plane: two or three Suzannes fill the view, the rest is off screen but in front */
static void scene_cull(int w, int h, int variant)
{
    scene_c4_setup(w, h, 8 | (6 << 8));
    glLoadIdentity();
    if (variant == 1) { glTranslatef(0.4f, -0.2f, -1.5f); glRotatef(50.0f, 0.0f, 1.0f, 0.0f); }
    else if (variant == 5) glTranslatef(0.9f, 0.4f, -1.6f);
    else glTranslatef(0.7f, -0.3f, -4.5f);
    if (variant == 2) {
        glViewport(w / 8, h / 6, w / 2, h / 2);
        glEnable(GL_SCISSOR_TEST);
        glScissor(w / 5, h / 5, w / 3, h / 3);
    }
    if (variant == 3) { glPolygonMode(GL_FRONT_AND_BACK, GL_LINE); glLineWidth(5.0f); }
    if (variant == 4) {
        glDisable(GL_LIGHTING); glDisable(GL_TEXTURE_2D);
        glDisableClientState(GL_NORMAL_ARRAY); glDisableClientState(GL_TEXTURE_COORD_ARRAY);
        glVertexPointer(2, GL_FLOAT, 32, (const void *)0);
        glColor3f(0.9f, 0.6f, 0.2f);
    }
    for (int frame = 0; frame < 2; frame++) {
        glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
        glDrawArrays(GL_TRIANGLES, 0, g_c4.nverts);
        if (frame == 0) glTranslatef(-0.25f, 0.1f, 0.0f);      /* the second frame is not the first one again */
    }
}

int scene_c4_vertex_count(void) { return g_c4.nverts; }
const void *scene_c4_host_data(void) { return g_c4.host; }
unsigned scene_c4_vbo(void) { return g_c4.vbo; }

/* ---------------------------------------------------------------- testbed-derived feature scenes */

/* 1.0-3: oversize triangle crossing every frustum plane, plus a quad poking through the near plane */
static void scene_clipping(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 20.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glShadeModel(variant ? GL_FLAT : GL_SMOOTH);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -3.0f);
    glRotatef(25.0f, 0.3f, 1.0f, 0.1f);
    glBegin(GL_TRIANGLES);
    glColor3f(1, 0, 0); glVertex3f(-9.0f, -4.0f, 0.5f);
    glColor3f(0, 1, 0); glVertex3f(9.0f, -5.0f, -1.0f);
    glColor3f(0, 0, 1); glVertex3f(0.5f, 8.0f, 2.9f);
    glColor3f(1, 1, 0); glVertex3f(-1.0f, -1.0f, 4.0f);
    glColor3f(0, 1, 1); glVertex3f(1.0f, -1.0f, -30.0f);
    glColor3f(1, 0, 1); glVertex3f(0.0f, 1.5f, 1.0f);
    glEnd();
    glBegin(GL_QUADS);
    glColor4f(1, 1, 1, 1);
    glVertex3f(-0.5f, -0.5f, 3.5f); glVertex3f(0.5f, -0.5f, 3.5f);
    glVertex3f(0.5f, 0.5f, -2.0f); glVertex3f(-0.5f, 0.5f, -2.0f);
    glEnd();
}

/* 1.0-8 / 1.0-4: every filled primitive mode, with and without back-face culling (1.0-5) */
static void scene_primitives(int w, int h, int variant)
{
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(-4.0, 4.0, -3.0, 3.0, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    glClearColor(0.1f, 0.1f, 0.1f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT);
    if (variant & 1) { glEnable(GL_CULL_FACE); glCullFace((variant & 2) ? GL_FRONT : GL_BACK); }
    if (variant & 4) glFrontFace(GL_CW);
    if (variant & 8) glShadeModel(GL_FLAT);
    static const GLenum modes[6] = { GL_TRIANGLES, GL_TRIANGLE_STRIP, GL_TRIANGLE_FAN, GL_QUADS, GL_QUAD_STRIP, GL_POLYGON };
    for (int m = 0; m < 6; m++) {
        glLoadIdentity();
        glTranslatef(-2.6f + 2.6f * (float)(m % 3), 1.4f - 2.8f * (float)(m / 3), 0.0f);
        glBegin(modes[m]);
        int n = (modes[m] == GL_TRIANGLES) ? 9 : (modes[m] == GL_QUADS) ? 8 : 7;
        for (int i = 0; i < n; i++) {
            float a = (float)i * 0.9f;
            glColor3f(0.5f + 0.5f * sinf(a), 0.5f + 0.5f * cosf(a * 1.3f), 0.5f + 0.5f * sinf(a * 0.7f + 1.0f));
            if (modes[m] == GL_TRIANGLE_STRIP || modes[m] == GL_QUAD_STRIP)
                glVertex2f(-1.0f + 0.33f * (float)i, (i & 1) ? 0.8f : -0.8f);
            else if (modes[m] == GL_TRIANGLE_FAN || modes[m] == GL_POLYGON)
                glVertex2f(i == 0 && modes[m] == GL_TRIANGLE_FAN ? 0.0f : 0.9f * cosf(a), i == 0 && modes[m] == GL_TRIANGLE_FAN ? 0.0f : 0.9f * sinf(a));
            else
                glVertex2f(-1.0f + 0.7f * (float)(i % 4) + 0.2f * (float)(i / 4), -0.8f + 0.55f * (float)((i * 5) % 4));
        }
        glEnd();
    }
}

/* 1.0-6 + depth-function sweep: two interpenetrating quads; variant = depth func index 0..7, bit 3 = no depth write,
 * bit 4 = shifted glDepthRange (exercises the double-precision depth expression) */
static void scene_zbuffer(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
    glClearDepth((variant & 7) >= 4 ? 0.0 : 1.0);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glDepthFunc(GL_NEVER + (variant & 7));
    if (variant & 8) glDepthMask(GL_FALSE);
    if (variant & 16) glDepthRange(0.25, 0.8);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -4.0f);
    for (int k = 0; k < 3; k++) {
        glPushMatrix();
        glRotatef(35.0f + 50.0f * (float)k, 0.2f, 1.0f, 0.1f * (float)k);
        glBegin(GL_QUADS);
        glColor3f(k == 0, k == 1, k == 2);
        glVertex3f(-1.2f, -1.0f, 0.0f); glVertex3f(1.2f, -1.0f, 0.0f);
        glColor3f(1.0f, 1.0f, k == 2);
        glVertex3f(1.2f, 1.0f, 0.0f); glVertex3f(-1.2f, 1.0f, 0.0f);
        glEnd();
        glPopMatrix();
    }
}

/* 1.0-9: rows of cubes in fog; variant 0 LINEAR, 1 EXP, 2 EXP2 */
static void scene_fog(int w, int h, int variant)
{
    static const GLint modes[3] = { GL_LINEAR, GL_EXP, GL_EXP2 };
    frustum_like_testbed(w, h, 50.0);
    GLfloat fog[] = { 0.5f, 0.5f, 0.5f, 0.3f };
    glClearColor(0.5f, 0.5f, 0.5f, 1.0f);
    glEnable(GL_DEPTH_TEST);
    glEnable(GL_FOG);
    glFogi(GL_FOG_MODE, modes[variant % 3]);
    glFogf(GL_FOG_START, 2.0f);
    glFogf(GL_FOG_END, 25.0f);
    glFogf(GL_FOG_DENSITY, 0.1f);
    glFogfv(GL_FOG_COLOR, fog);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    for (int row = 0; row < 5; row++)
        for (int col = -2; col <= 2; col++) {
            glLoadIdentity();
            glTranslatef(2.2f * (float)col, -1.0f, -4.0f - 4.5f * (float)row);
            glRotatef(20.0f * (float)(row + col), 0.0f, 1.0f, 0.0f);
            glBegin(GL_QUADS);
            glColor3f(0.9f, 0.3f + 0.1f * (float)row, 0.2f);
            glVertex3f(-0.7f, -0.7f, 0.7f); glVertex3f(0.7f, -0.7f, 0.7f); glVertex3f(0.7f, 0.7f, 0.7f); glVertex3f(-0.7f, 0.7f, 0.7f);
            glColor3f(0.2f, 0.8f, 0.3f);
            glVertex3f(0.7f, -0.7f, 0.7f); glVertex3f(0.7f, -0.7f, -0.7f); glVertex3f(0.7f, 0.7f, -0.7f); glVertex3f(0.7f, 0.7f, 0.7f);
            glColor3f(0.2f, 0.3f, 0.9f);
            glVertex3f(-0.7f, 0.7f, 0.7f); glVertex3f(0.7f, 0.7f, 0.7f); glVertex3f(0.7f, 0.7f, -0.7f); glVertex3f(-0.7f, 0.7f, -0.7f);
            glEnd();
        }
}

/* 1.0-11: overlapping translucent quads; variant selects the blend function pair */
static void scene_blend(int w, int h, int variant)
{
    static const GLenum src[8] = { GL_SRC_ALPHA, GL_SRC_ALPHA, GL_ONE, GL_DST_COLOR, GL_ONE_MINUS_DST_ALPHA, GL_SRC_ALPHA_SATURATE, GL_ONE_MINUS_SRC_COLOR, GL_ZERO };
    static const GLenum dst[8] = { GL_ONE_MINUS_SRC_ALPHA, GL_ONE, GL_ONE, GL_ZERO, GL_DST_ALPHA, GL_ONE, GL_SRC_COLOR, GL_ONE_MINUS_DST_COLOR };
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(-2.0, 2.0, -1.5, 1.5, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    glClearColor(0.3f, 0.3f, 0.3f, 0.5f);
    glClear(GL_COLOR_BUFFER_BIT);
    glEnable(GL_BLEND);
    glBlendFunc(src[variant & 7], dst[variant & 7]);
    for (int k = 0; k < 4; k++) {
        glLoadIdentity();
        glTranslatef(-0.9f + 0.6f * (float)k, -0.3f + 0.2f * (float)k, 0.0f);
        glRotatef(13.0f * (float)k, 0, 0, 1);
        glBegin(GL_QUADS);
        glColor4f(k == 0 || k == 3, k == 1 || k == 3, k == 2, 0.35f + 0.15f * (float)k);
        glVertex2f(-0.8f, -0.8f); glVertex2f(0.8f, -0.8f); glVertex2f(0.8f, 0.8f); glVertex2f(-0.8f, 0.8f);
        glEnd();
    }
}

/* 1.0-17: stencil mask pass then GL_EQUAL pass; variant perturbs ops / masks */
static void scene_stencil(int w, int h, int variant)
{
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(-1.5, 1.5, -1.0, 1.0, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    glClearColor(0.2f, 0.2f, 0.3f, 1.0f);
    glClearStencil(variant == 3 ? 200 : 0);
    glClear(GL_COLOR_BUFFER_BIT | GL_STENCIL_BUFFER_BIT);
    glEnable(GL_STENCIL_TEST);
    glStencilFunc(GL_ALWAYS, 1, 0xFF);
    glStencilOp(GL_KEEP, GL_KEEP, variant == 1 ? GL_INCR : variant == 2 ? GL_INVERT : variant == 3 ? GL_DECR_WRAP : GL_REPLACE);
    if (variant == 2) glStencilMask(0x0F);
    glColorMask(GL_FALSE, GL_FALSE, GL_FALSE, GL_FALSE);
    glBegin(GL_TRIANGLE_FAN);
    glVertex2f(0.0f, 0.0f);
    for (int i = 0; i <= 32; i++) {
        float a = (float)i * 2.0f * 3.14159f / 32.0f;
        glVertex2f(0.5f * cosf(a), 0.5f * sinf(a));
    }
    glEnd();
    glStencilFunc(variant == 1 ? GL_LEQUAL : GL_EQUAL, variant == 2 ? 0x0F : variant == 3 ? 199 : 1, variant == 2 ? 0x0F : 0xFF);
    glStencilOp(variant == 1 ? GL_ZERO : GL_KEEP, GL_KEEP, variant == 1 ? GL_DECR : GL_KEEP);
    glColorMask(GL_TRUE, GL_TRUE, variant == 3 ? GL_FALSE : GL_TRUE, GL_TRUE);
    glLoadIdentity();
    glRotatef(31.0f, 0.0f, 0.0f, 1.0f);
    glBegin(GL_QUADS);
    glColor3f(1.0f, 0.0f, 0.0f); glVertex2f(-0.8f, -0.8f);
    glColor3f(0.0f, 1.0f, 0.0f); glVertex2f(0.8f, -0.8f);
    glColor3f(0.0f, 0.0f, 1.0f); glVertex2f(0.8f, 0.8f);
    glColor3f(1.0f, 1.0f, 0.0f); glVertex2f(-0.8f, 0.8f);
    glEnd();
    glDisable(GL_STENCIL_TEST);
}

/* 1.0-10: lit sphere from triangle strips with GL_NORMALIZE; variant 0 flat, 1 smooth, 2 phong,
 * 3 smooth + spotlight + attenuation, 4 smooth + COLOR_MATERIAL, 5 phong + spotlight */
static void scene_lighting(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glEnable(GL_NORMALIZE);
    glClearColor(0.05f, 0.05f, 0.1f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glEnable(GL_LIGHTING);
    glEnable(GL_LIGHT0);
    GLfloat pos[] = { 1.5f, 2.0f, 1.0f, 1.0f }, dif[] = { 1.0f, 0.9f, 0.8f, 1.0f }, spc[] = { 1, 1, 1, 1 }, la[] = { 0.1f, 0.1f, 0.1f, 1 };
    glLightfv(GL_LIGHT0, GL_DIFFUSE, dif);
    glLightfv(GL_LIGHT0, GL_SPECULAR, spc);
    glLightfv(GL_LIGHT0, GL_AMBIENT, la);
    GLfloat md[] = { 0.2f, 0.5f, 0.9f, 1.0f }, ms[] = { 0.8f, 0.8f, 0.8f, 1.0f };
    glMaterialfv(GL_FRONT, GL_DIFFUSE, md);
    glMaterialfv(GL_FRONT, GL_SPECULAR, ms);
    glMaterialf(GL_FRONT, GL_SHININESS, 48.0f);
    glShadeModel(variant == 0 ? GL_FLAT : (variant == 2 || variant == 5) ? GL_PHONG : GL_SMOOTH);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -3.0f);
    glLightfv(GL_LIGHT0, GL_POSITION, pos);
    if (variant == 3 || variant == 5) {
        GLfloat dir[] = { -0.45f, -0.6f, -0.66f };
        glLightfv(GL_LIGHT0, GL_SPOT_DIRECTION, dir);
        glLightf(GL_LIGHT0, GL_SPOT_CUTOFF, 25.0f);
        glLightf(GL_LIGHT0, GL_SPOT_EXPONENT, 6.0f);
        glLightf(GL_LIGHT0, GL_CONSTANT_ATTENUATION, 0.5f);
        glLightf(GL_LIGHT0, GL_LINEAR_ATTENUATION, 0.1f);
        glLightf(GL_LIGHT0, GL_QUADRATIC_ATTENUATION, 0.02f);
    }
    if (variant == 4) { glEnable(GL_COLOR_MATERIAL); glColorMaterial(GL_FRONT, GL_DIFFUSE); }
    glScalef(1.0f, 1.2f, 0.9f);
    const int stacks = 18, slices = 24;
    for (int i = 0; i < stacks; i++) {
        float t0 = 3.14159265f * (float)i / stacks, t1 = 3.14159265f * (float)(i + 1) / stacks;
        glBegin(GL_TRIANGLE_STRIP);
        for (int j = 0; j <= slices; j++) {
            float p = 2.0f * 3.14159265f * (float)j / slices;
            float x0 = sinf(t0) * cosf(p), y0 = cosf(t0), z0 = sinf(t0) * sinf(p);
            float x1 = sinf(t1) * cosf(p), y1 = cosf(t1), z1 = sinf(t1) * sinf(p);
            if (variant == 4) glColor3f(0.5f + 0.5f * x0, 0.5f + 0.5f * y0, 0.5f + 0.5f * z0);
            glNormal3f(2.0f * x0, 2.0f * y0, 2.0f * z0); glVertex3f(x0, y0, z0);
            glNormal3f(2.0f * x1, 2.0f * y1, 2.0f * z1); glVertex3f(x1, y1, z1);
        }
        glEnd();
    }
}

/* 1.0-13: the same geometry through a compiled display list, called with different matrices;
 * variant 1 uses GL_COMPILE_AND_EXECUTE and a nested list */
static void scene_displaylist(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -6.0f);
    GLuint base = glGenLists(2);
    glNewList(base + 1, GL_COMPILE);
    glBegin(GL_TRIANGLES);
    glColor3f(1, 0, 0); glVertex3f(-0.5f, -0.5f, 0.0f);
    glColor3f(0, 1, 0); glVertex3f(0.5f, -0.5f, 0.0f);
    glColor3f(0, 0, 1); glVertex3f(0.0f, 0.6f, 0.0f);
    glEnd();
    glEndList();
    glNewList(base, variant ? GL_COMPILE_AND_EXECUTE : GL_COMPILE);
    glPushMatrix();
    glRotatef(30.0f, 0, 0, 1);
    glCallList(base + 1);
    glTranslatef(0.0f, 0.0f, 0.5f);
    glScalef(0.5f, 0.5f, 1.0f);
    glCallList(base + 1);
    glPopMatrix();
    glEndList();
    for (int k = 0; k < 5; k++) {
        glPushMatrix();
        glTranslatef(-2.4f + 1.2f * (float)k, 0.3f * (float)(k % 2), -0.4f * (float)k);
        glCallList(base);
        glPopMatrix();
    }
    glDeleteLists(base, 2);
}

/* 1.5-0: interleaved position+colour VBO drawn as quads (stride 24), plus an indexed draw with
 * ubyte colours, a client-memory array draw and an element-buffer draw */
static void scene_vbo(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.1f, 0.1f, 0.15f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    static const float cube[24][6] = {
        { -1, -1, 1, 1, 0, 0 }, { 1, -1, 1, 1, 0, 0 }, { 1, 1, 1, 1, 0.5f, 0 }, { -1, 1, 1, 1, 0.5f, 0 },
        { -1, -1, -1, 0, 1, 0 }, { -1, 1, -1, 0, 1, 0 }, { 1, 1, -1, 0, 1, 0.5f }, { 1, -1, -1, 0, 1, 0.5f },
        { -1, 1, -1, 0, 0, 1 }, { -1, 1, 1, 0, 0, 1 }, { 1, 1, 1, 0.5f, 0, 1 }, { 1, 1, -1, 0.5f, 0, 1 },
        { -1, -1, -1, 1, 1, 0 }, { 1, -1, -1, 1, 1, 0 }, { 1, -1, 1, 1, 1, 0.5f }, { -1, -1, 1, 1, 1, 0.5f },
        { 1, -1, -1, 1, 0, 1 }, { 1, 1, -1, 1, 0, 1 }, { 1, 1, 1, 1, 0.5f, 1 }, { 1, -1, 1, 1, 0.5f, 1 },
        { -1, -1, -1, 0, 1, 1 }, { -1, -1, 1, 0, 1, 1 }, { -1, 1, 1, 0.5f, 1, 1 }, { -1, 1, -1, 0.5f, 1, 1 },
    };
    GLuint vbo[3];
    glGenBuffers(3, vbo);
    glBindBuffer(GL_ARRAY_BUFFER, vbo[0]);
    glBufferData(GL_ARRAY_BUFFER, sizeof cube, cube, GL_STATIC_DRAW);
    glEnableClientState(GL_VERTEX_ARRAY);
    glEnableClientState(GL_COLOR_ARRAY);
    glVertexPointer(3, GL_FLOAT, 24, (const void *)0);
    glColorPointer(3, GL_FLOAT, 24, (const void *)12);
    glLoadIdentity();
    glTranslatef(-1.6f, 0.0f, -6.0f);
    glRotatef(35.0f, 1.0f, 1.0f, 0.0f);
    glDrawArrays(GL_QUADS, 0, 24);

    /* indexed, ubyte colours with 4 components, positions of size 2, element buffer with ushort indices */
    struct { float x, y; uint8_t c[4]; } flat[5] = {
        { -0.8f, -0.8f, { 255, 0, 0, 255 } }, { 0.8f, -0.8f, { 0, 255, 0, 255 } }, { 0.8f, 0.8f, { 0, 0, 255, 128 } },
        { -0.8f, 0.8f, { 255, 255, 0, 255 } }, { 0.0f, 1.4f, { 255, 255, 255, 64 } } };
    uint16_t idx[9] = { 0, 1, 2, 0, 2, 3, 3, 2, 4 };
    glBindBuffer(GL_ARRAY_BUFFER, vbo[1]);
    glBufferData(GL_ARRAY_BUFFER, sizeof flat, flat, GL_STATIC_DRAW);
    glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, vbo[2]);
    glBufferData(GL_ELEMENT_ARRAY_BUFFER, sizeof idx, idx, GL_STATIC_DRAW);
    glVertexPointer(2, GL_FLOAT, 12, (const void *)0);
    glColorPointer(4, GL_UNSIGNED_BYTE, 12, (const void *)8);
    glLoadIdentity();
    glTranslatef(1.5f, -0.3f, -5.0f);
    if (variant & 1) glEnable(GL_BLEND), glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA);
    glDrawElements(GL_TRIANGLES, 9, GL_UNSIGNED_SHORT, (const void *)0);
    glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, 0);
    /* client-memory indices against the bound array buffer */
    uint8_t idx8[3] = { 4, 3, 0 };
    glLoadIdentity();
    glTranslatef(1.5f, 0.9f, -5.5f);
    glDrawElements(GL_TRIANGLES, 3, GL_UNSIGNED_BYTE, idx8);

    /* client-memory arrays (no buffer bound): snapshotted at call time */
    glBindBuffer(GL_ARRAY_BUFFER, 0);
    float tri[3][7] = { { 0, 0, 0, 1, 0, 1, 1 }, { 1, 0, 0, 0, 1, 1, 1 }, { 0, 1, 0, 1, 1, 0, 1 } };
    glVertexPointer(3, GL_FLOAT, 28, &tri[0][0]);
    glColorPointer(4, GL_FLOAT, 28, &tri[0][3]);
    glLoadIdentity();
    glTranslatef(-0.4f, -1.7f, -4.0f);
    glDrawArrays(GL_TRIANGLES, 0, 3);
    tri[0][0] = 99.0f; /* must not affect the draw above */
    glDisableClientState(GL_COLOR_ARRAY);
    glDisableClientState(GL_VERTEX_ARRAY);
    glDeleteBuffers(3, vbo);
}

/* 1.0-16: hostile input must not crash and must match the reference: NaN/Inf colours,
 * huge coordinates, degenerate triangles */
static void scene_validation(int w, int h, int variant)
{
    (void)variant;
    frustum_like_testbed(w, h, 100.0);
    glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glEnable(GL_DEPTH_TEST);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -3.0f);
    float inf = 1.0f / 0.0f * 1.0f, nanv = inf - inf;
    glBegin(GL_TRIANGLES);
    glColor3f(nanv, 0.5f, inf); glVertex3f(-1.0f, -1.0f, 0.0f);
    glColor4f(2.0f, -1.0f, 0.5f, nanv); glVertex3f(1.0f, -1.0f, 0.0f);
    glColor3f(0.2f, 0.9f, 0.4f); glVertex3f(0.0f, 1.0f, 0.0f);
    /* far outside the frustum */
    glColor3f(1, 1, 1);
    glVertex3f(-1e10f, -1e10f, -5.0f); glVertex3f(1e10f, -1e10f, -5.0f); glVertex3f(0.0f, 1e10f, -5.0f);
    /* degenerate: zero area, repeated vertices */
    glVertex3f(0.3f, 0.3f, 0.0f); glVertex3f(0.3f, 0.3f, 0.0f); glVertex3f(0.3f, 0.3f, 0.0f);
    glVertex3f(-0.5f, 0.0f, 0.0f); glVertex3f(0.0f, 0.0f, 0.0f); glVertex3f(0.5f, 0.0f, 0.0f);
    /* crossing w = 0 */
    glColor3f(0.8f, 0.2f, 0.2f);
    glVertex3f(-0.5f, -0.5f, 2.0f); glVertex3f(0.5f, -0.5f, 4.0f); glVertex3f(0.0f, 0.5f, 3.0f);
    glEnd();
}

/* scissor + partial clears + colour mask + viewport offset (Appendix A.22, A.24) */
static void scene_scissor(int w, int h, int variant)
{
    glClearColor(0.1f, 0.2f, 0.3f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT | GL_STENCIL_BUFFER_BIT);
    glEnable(GL_SCISSOR_TEST);
    glScissor(w / 8, h / 6, w / 2, h / 2);
    glClearColor(0.6f, 0.1f, 0.1f, 1.0f);
    glColorMask(GL_FALSE, GL_TRUE, GL_TRUE, GL_TRUE); /* glClear ignores the colour mask */
    glClear(GL_COLOR_BUFFER_BIT);
    glColorMask(GL_TRUE, GL_TRUE, GL_TRUE, GL_TRUE);
    glViewport(w / 10, h / 8, (w * 3) / 4, (h * 2) / 3);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    if (variant & 1) glScissor(-5, -7, w / 3, h + 50);
    if (variant & 2) glColorMask(GL_TRUE, GL_FALSE, GL_TRUE, GL_FALSE);
    glBegin(GL_TRIANGLES);
    glColor3f(1, 1, 0); glVertex2f(-1.2f, -1.1f);
    glColor3f(0, 1, 1); glVertex2f(1.3f, -0.9f);
    glColor3f(1, 0, 1); glVertex2f(0.1f, 1.4f);
    glEnd();
    glDisable(GL_SCISSOR_TEST);
}

/* textures: non-square LA texture, clamp vs repeat, nearest vs bilinear (1.0-12), texture matrix,
 * affine hint, alpha test without blending */
static void scene_texture_misc(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glClearColor(0.0f, 0.1f, 0.0f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glEnable(GL_TEXTURE_2D);
    make_gradient_la(48, 20);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, (variant & 1) ? GL_LINEAR_MIPMAP_NEAREST : GL_NEAREST_MIPMAP_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, (variant & 1) ? GL_LINEAR : GL_NEAREST);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, (variant & 2) ? GL_CLAMP_TO_EDGE : GL_REPEAT);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, (variant & 4) ? GL_CLAMP : GL_REPEAT);
    if (variant & 8) glHint(GL_PERSPECTIVE_CORRECTION_HINT, GL_FASTEST);
    glEnable(GL_ALPHA_TEST);
    glAlphaFunc(GL_GEQUAL, 0.5f);
    glMatrixMode(GL_TEXTURE);
    glLoadIdentity();
    glTranslatef(0.25f, -0.1f, 0.0f);
    glRotatef(15.0f, 0, 0, 1);
    glScalef(1.5f, 0.75f, 1.0f);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -2.5f);
    glRotatef(-50.0f, 1.0f, 0.2f, 0.0f);
    glBegin(GL_QUADS);
    glColor3f(1.0f, 0.8f, 0.8f); glTexCoord2f(-1.0f, -1.0f); glVertex3f(-1.5f, -1.5f, 0.0f);
    glColor3f(0.8f, 1.0f, 0.8f); glTexCoord2f(2.0f, -1.0f); glVertex3f(1.5f, -1.5f, 0.0f);
    glColor3f(0.8f, 0.8f, 1.0f); glTexCoord2f(2.0f, 2.0f); glVertex3f(1.5f, 1.5f, 0.0f);
    glColor3f(1.0f, 1.0f, 1.0f); glTexCoord2f(-1.0f, 2.0f); glVertex3f(-1.5f, 1.5f, 0.0f);
    glEnd();
}

/* many small state changes inside one frame: exercises state-block batching and ordering */
static void scene_state_churn(int w, int h, int variant)
{
    (void)variant;
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(0.0, 16.0, 0.0, 12.0, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glClearColor(0, 0, 0, 1);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT | GL_STENCIL_BUFFER_BIT);
    GLuint tex = make_checker_rgb(16, 2, 0);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
    for (int k = 0; k < 48; k++) {
        glLoadIdentity();
        glTranslatef((float)(k % 8) * 1.9f + 0.4f, (float)(k / 8) * 1.9f + 0.3f, 0.0f);
        if (k % 3 == 0) glEnable(GL_BLEND); else glDisable(GL_BLEND);
        glBlendFunc(GL_SRC_ALPHA, (k % 2) ? GL_ONE : GL_ONE_MINUS_SRC_ALPHA);
        if (k % 4 == 1) { glEnable(GL_TEXTURE_2D); glBindTexture(GL_TEXTURE_2D, tex); } else glDisable(GL_TEXTURE_2D);
        if (k % 5 == 2) glShadeModel(GL_FLAT); else glShadeModel(GL_SMOOTH);
        if (k == 20) glClear(GL_DEPTH_BUFFER_BIT);     /* a clear in the middle of the frame */
        glBegin(GL_TRIANGLE_FAN);
        glColor4f(1.0f, (float)(k % 7) / 6.0f, 0.2f, 0.6f);
        glTexCoord2f(0.5f, 0.5f); glVertex2f(1.0f, 1.0f);
        for (int i = 0; i <= 6; i++) {
            float a = 6.2831853f * (float)i / 6.0f;
            if (i == 3) glColor4f(0.1f, 0.3f, 1.0f, 0.9f);  /* colour change between vertices */
            glTexCoord2f(0.5f + 0.5f * cosf(a), 0.5f + 0.5f * sinf(a));
            glVertex2f(1.0f + 1.3f * cosf(a), 1.0f + 1.3f * sinf(a));
        }
        glEnd();
    }
}

/* 1.0-1 / 1.0-8: points, lines, line strips and loops; variant bit0 wide lines + big points, bit1 depth test,
 * bit2 blending, bit3 fog (EXP2: the only mode that visibly fogs lines), bit4 textured (no alpha test:
 * the reference's line loop never terminates when the alpha test rejects a pixel), bit5 scissor */
static void scene_lines(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 50.0);
    glClearColor(0.05f, 0.05f, 0.08f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    if (variant & 1) { glLineWidth(3.0f); glPointSize(5.0f); }
    if (variant & 2) glEnable(GL_DEPTH_TEST);
    if (variant & 4) { glEnable(GL_BLEND); glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA); }
    if (variant & 8) {
        GLfloat fog[] = { 0.4f, 0.1f, 0.1f, 0.6f };
        glEnable(GL_FOG); glFogi(GL_FOG_MODE, GL_EXP2); glFogf(GL_FOG_DENSITY, 0.25f); glFogfv(GL_FOG_COLOR, fog);
    }
    if (variant & 16) {
        glEnable(GL_TEXTURE_2D);
        make_checker_rgb(16, 2, 1);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    }
    if (variant & 32) { glEnable(GL_SCISSOR_TEST); glScissor(w / 5, h / 7, w / 2, (h * 2) / 3); }
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -3.0f);
    glRotatef(25.0f, 0.0f, 0.0f, 1.0f);
    glRotatef(35.0f, 1.0f, 0.0f, 0.0f);
    glBegin(GL_LINES);
    for (int i = 0; i < 24; i++) {
        float a = 6.2831853f * (float)i / 24.0f;
        glColor4f(0.5f + 0.5f * cosf(a), 0.5f + 0.5f * sinf(a), 0.6f, 0.7f);
        glTexCoord2f(0.0f, 0.0f); glVertex3f(0.1f * cosf(a), 0.1f * sinf(a), 0.0f);
        glColor4f(1.0f, 1.0f, 0.2f, 0.4f);
        glTexCoord2f(3.0f, 1.0f); glVertex3f(2.6f * cosf(a), 2.6f * sinf(a), (float)(i % 5) * 0.6f - 1.0f);    /* some leave the frustum */
    }
    glEnd();
    glBegin(GL_LINE_STRIP);
    for (int i = 0; i < 40; i++) {
        float a = (float)i * 0.37f;
        glColor3f(0.2f, 0.9f, 0.5f + 0.5f * sinf(a));
        glTexCoord2f((float)i * 0.1f, 0.5f);
        glVertex3f(-1.5f + 0.075f * (float)i, 0.8f * sinf(a), 0.5f * cosf(a));
    }
    glEnd();
    glBegin(GL_LINE_LOOP);
    for (int i = 0; i < 7; i++) {
        float a = 6.2831853f * (float)i / 7.0f;
        glColor3f(1.0f, 0.3f, 0.3f);
        glVertex3f(1.1f * cosf(a), 1.1f * sinf(a), -0.4f);
    }
    glEnd();
    glBegin(GL_POINTS);
    for (int i = 0; i < 150; i++) {
        float a = (float)i * 0.61f;
        glColor4f(0.3f + 0.7f * fabsf(sinf(a)), 0.8f, 0.3f + 0.7f * fabsf(cosf(a)), 0.8f);
        glTexCoord2f(a, a * 0.3f);
        glVertex3f(1.9f * cosf(a) * (float)(i % 10) / 9.0f, 1.4f * sinf(a * 1.7f), 1.5f * sinf(a * 0.3f));
    }
    glEnd();
    /* degenerate and axis-aligned segments */
    glBegin(GL_LINES);
    glColor3f(1, 1, 1);
    glVertex3f(0.5f, 0.5f, 0.0f); glVertex3f(0.5f, 0.5f, 0.0f);
    glVertex3f(-1.0f, -0.9f, 0.0f); glVertex3f(1.0f, -0.9f, 0.0f);
    glVertex3f(-1.2f, -1.0f, 0.0f); glVertex3f(-1.2f, 1.0f, 0.0f);
    glEnd();
}

/* polygon modes: Suzanne as wireframe (the testbed's SPACE toggle, 1.0-18) or points; variant bit0 GL_POINT instead of
 * GL_LINE, bit1 GL_PHONG (colours re-lit per vertex, raster.c:864-868), bit2 back faces filled / front outlined,
 * bit3 cull back faces */
static void scene_wireframe(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.2f, 0.2f, 0.3f, 1.0f);
    suzanne_lights_and_material();
    glShadeModel((variant & 2) ? GL_PHONG : GL_SMOOTH);
    GLenum mode = (variant & 1) ? GL_POINT : GL_LINE;
    if (variant & 4) { glPolygonMode(GL_FRONT, mode); glPolygonMode(GL_BACK, GL_FILL); }
    else glPolygonMode(GL_FRONT_AND_BACK, mode);
    if (variant & 8) glEnable(GL_CULL_FACE);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -2.2f);
    glRotatef(20.0f, 1.0f, 0.0f, 0.0f);
    glRotatef(30.0f, 0.0f, 1.0f, 0.0f);
    suzanne_immediate();
    /* an oversize outlined triangle: the outline follows the CLIPPED polygon's fan (raster.c:916-945) */
    glDisable(GL_LIGHTING);
    glBegin(GL_TRIANGLES);
    glColor3f(1, 0, 0); glVertex3f(-6.0f, -2.0f, 1.0f);
    glColor3f(0, 1, 0); glVertex3f(6.0f, -2.5f, 0.0f);
    glColor3f(0, 0, 1); glVertex3f(0.3f, 5.0f, 1.9f);
    glEnd();
}

/* Depth ordering stress for the order-independent visibility path (k_vis.cu): exact depth ties between coplanar
 * duplicates, all four ordering depth functions, non-default depth range, a mid-frame change of depth function,
 * more than a thousand tiny triangles inside one 64x64 tile, large and small triangles mixed, scissor.
 *   variant & 3   : GL_LESS, GL_LEQUAL, GL_GREATER, GL_GEQUAL
 *   variant & 4   : glDepthRange(0.3, 0.9)
 *   variant & 8   : second half of the frame uses the "opposite tie" function (LESS<->LEQUAL, GREATER<->GEQUAL)
 *   variant & 16  : scissor rectangle */
static void scene_depth_order(int w, int h, int variant)
{
    static const GLenum funcs[4] = { GL_LESS, GL_LEQUAL, GL_GREATER, GL_GEQUAL };
    const int fi = variant & 3;
    frustum_like_testbed(w, h, 60.0);
    glEnable(GL_DEPTH_TEST);
    glDepthFunc(funcs[fi]);
    if (variant & 4) glDepthRange(0.3, 0.9);
    glClearColor(0.1f, 0.1f, 0.15f, 1.0f);
    glClearDepth(fi >= 2 ? 0.0 : 1.0);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    if (variant & 16) { glEnable(GL_SCISSOR_TEST); glScissor(w / 9, h / 7, (w * 3) / 4, (h * 2) / 3); }
    glLoadIdentity();
    glTranslatef(0.0f, 0.0f, -5.0f);

    /* 1. a large tilted background quad and an interpenetrating one */
    glBegin(GL_QUADS);
    glColor3f(0.8f, 0.2f, 0.2f);
    glVertex3f(-4.0f, -3.0f, -2.0f); glVertex3f(4.0f, -3.0f, 1.0f); glVertex3f(4.0f, 3.0f, 1.0f); glVertex3f(-4.0f, 3.0f, -2.0f);
    glColor3f(0.2f, 0.8f, 0.2f);
    glVertex3f(-4.0f, -3.0f, 1.0f); glVertex3f(4.0f, -3.0f, -2.0f); glVertex3f(4.0f, 3.0f, -2.0f); glVertex3f(-4.0f, 3.0f, 1.0f);
    glEnd();

    /* 2. exact ties: the same triangle five times with different colours, then shifted coplanar copies that share
     *    vertices with it (identical vertex data -> identical interpolated depth on the shared pixels) */
    static const GLfloat tri[3][3] = { { -1.5f, -1.0f, 1.5f }, { 1.2f, -0.8f, 1.5f }, { 0.1f, 1.4f, 1.5f } };
    for (int k = 0; k < 5; k++) {
        glBegin(GL_TRIANGLES);
        glColor3f(0.2f * (float)k, 1.0f - 0.2f * (float)k, 0.5f);
        for (int j = 0; j < 3; j++) glVertex3f(tri[j][0], tri[j][1], tri[j][2]);
        glEnd();
    }
    glBegin(GL_TRIANGLES);
    glColor3f(1.0f, 1.0f, 0.0f);
    glVertex3f(tri[1][0], tri[1][1], tri[1][2]); glVertex3f(2.6f, 1.3f, 1.5f); glVertex3f(tri[2][0], tri[2][1], tri[2][2]);
    glColor3f(0.0f, 1.0f, 1.0f);
    glVertex3f(tri[2][0], tri[2][1], tri[2][2]); glVertex3f(2.6f, 1.3f, 1.5f); glVertex3f(tri[1][0], tri[1][1], tri[1][2]);       /* the same triangle, other winding */
    glEnd();

    if (variant & 8) glDepthFunc(funcs[fi ^ 1]);

    /* 3. a dense patch: 36 x 36 cells of two tiny triangles each, inside roughly one tile, drawn twice (ties again) */
    for (int pass = 0; pass < 2; pass++) {
        glBegin(GL_TRIANGLES);
        for (int iy = 0; iy < 36; iy++)
            for (int ix = 0; ix < 36; ix++) {
                const float x0 = -2.4f + 0.035f * (float)ix, y0 = -1.6f + 0.035f * (float)iy, d = 0.035f;
                const float z = 1.8f + 0.01f * (float)((ix * 7 + iy * 3) % 5);
                glColor3f((float)((ix + pass) & 1), (float)(iy & 1), pass ? 0.9f : 0.1f);
                glVertex3f(x0, y0, z); glVertex3f(x0 + d, y0, z); glVertex3f(x0 + d, y0 + d, z);
                glVertex3f(x0, y0, z); glVertex3f(x0 + d, y0 + d, z); glVertex3f(x0, y0 + d, z);
            }
        glEnd();
    }

    /* 4. long thin slivers crossing many tiles, and a strip that wraps back over itself */
    glBegin(GL_TRIANGLES);
    for (int k = 0; k < 12; k++) {
        const float y = -2.5f + 0.4f * (float)k;
        glColor3f(0.3f + 0.05f * (float)k, 0.3f, 1.0f - 0.05f * (float)k);
        glVertex3f(-3.8f, y, 0.5f + 0.1f * (float)k); glVertex3f(3.8f, y + 0.05f, 2.0f - 0.1f * (float)k); glVertex3f(3.8f, y + 0.12f, 0.7f);
    }
    glEnd();
    glBegin(GL_TRIANGLE_STRIP);
    for (int k = 0; k < 20; k++) {
        const float a = 0.45f * (float)k;
        glColor3f(0.5f + 0.5f * sinf(a), 0.5f + 0.5f * cosf(a), 0.6f);
        glVertex3f(1.5f + 1.2f * cosf(a), -0.5f + 1.2f * sinf(a), 1.0f + 0.05f * (float)k);
        glVertex3f(1.5f + 0.6f * cosf(a), -0.5f + 0.6f * sinf(a), 2.2f - 0.05f * (float)k);
    }
    glEnd();
}

/* A buffer too large for a host mirror (> 8 MB): the "current" normal / colour an array draw leaves behind is the
 * last element's (gl_api.c:1826-1842) and must follow glBufferSubData / glBufferData without the host re-reading HBM.
 * Each step draws the (mostly degenerate) array, then an immediate-mode lit triangle that uses the current normal
 * and a flat one that uses the current colour. */
static void scene_vbo_large(int w, int h, int variant)
{
    (void)variant;
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.1f, 0.1f, 0.15f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    const int n = 240000;                                      /* 240000 x 40 B = 9.6 MB */
    float *v = (float *)calloc((size_t)n * 10, sizeof(float)); /* pos3 normal3 colour4 */
    for (int i = 0; i < n; i++) { v[i * 10 + 5] = 1.0f; v[i * 10 + 6] = 0.5f; v[i * 10 + 7] = 0.5f; v[i * 10 + 8] = 0.5f; v[i * 10 + 9] = 1.0f; }
    for (int i = 0; i < 300; i++) {                            /* a fan of real triangles in front */
        float a = 0.0209f * (float)i;
        v[i * 10] = (i % 3 == 0) ? 0.0f : 1.2f * cosf(a); v[i * 10 + 1] = (i % 3 == 0) ? 0.0f : 1.2f * sinf(a); v[i * 10 + 2] = 0.0f;
        v[i * 10 + 6] = 0.2f + 0.002f * (float)i;
    }
    GLuint vbo;
    glGenBuffers(1, &vbo);
    glBindBuffer(GL_ARRAY_BUFFER, vbo);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)((size_t)n * 40), v, GL_DYNAMIC_DRAW);
    glEnableClientState(GL_VERTEX_ARRAY);
    glEnableClientState(GL_NORMAL_ARRAY);
    glEnableClientState(GL_COLOR_ARRAY);
    glVertexPointer(3, GL_FLOAT, 40, (const void *)0);
    glNormalPointer(GL_FLOAT, 40, (const void *)12);
    glColorPointer(4, GL_FLOAT, 40, (const void *)24);
    GLfloat lpos[4] = { 0.3f, 0.5f, 1.0f, 0.0f }, ldif[4] = { 0.9f, 0.8f, 0.7f, 1.0f };
    glLightfv(GL_LIGHT0, GL_POSITION, lpos);
    glLightfv(GL_LIGHT0, GL_DIFFUSE, ldif);
    glEnable(GL_LIGHT0);
    for (int step = 0; step < 4; step++) {
        float last[7] = { 0.0f, 0.0f, 1.0f, 0.5f, 0.5f, 0.5f, 1.0f };
        if (step == 1) { last[0] = 0.6f; last[2] = 0.8f; last[3] = 0.9f; last[4] = 0.2f; }
        if (step == 2) { last[1] = -0.6f; last[2] = 0.8f; last[3] = 0.1f; last[5] = 0.9f; }
        if (step == 1) glBufferSubData(GL_ARRAY_BUFFER, (GLintptr)((size_t)(n - 1) * 40 + 12), 28, last);
        if (step == 2) {                                       /* partial overlap: normal.z and the colour only */
            glBufferSubData(GL_ARRAY_BUFFER, (GLintptr)((size_t)(n - 1) * 40 + 20), 20, last + 2);
        }
        if (step == 3) {                                       /* fresh contents through glBufferData */
            v[(size_t)(n - 1) * 10 + 3] = -0.7f; v[(size_t)(n - 1) * 10 + 4] = 0.1f; v[(size_t)(n - 1) * 10 + 5] = 0.7f;
            v[(size_t)(n - 1) * 10 + 6] = 0.3f; v[(size_t)(n - 1) * 10 + 7] = 0.9f; v[(size_t)(n - 1) * 10 + 8] = 0.4f;
            glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)((size_t)n * 40), v, GL_DYNAMIC_DRAW);
        }
        glDisable(GL_LIGHTING);
        glLoadIdentity();
        glTranslatef(-2.4f + 1.6f * (float)step, 1.2f, -7.0f);
        glDrawArrays(GL_TRIANGLES, 0, n);
        const float x = -2.4f + 1.6f * (float)step;
        glLoadIdentity();
        glTranslatef(x, -0.9f, -6.0f);
        glBegin(GL_TRIANGLES);                                 /* flat: the current colour */
        glVertex3f(-0.6f, -0.5f, 0.0f); glVertex3f(0.0f, -0.5f, 0.0f); glVertex3f(-0.3f, 0.4f, 0.0f);
        glEnd();
        glEnable(GL_LIGHTING);
        glBegin(GL_TRIANGLES);                                 /* lit: the current normal */
        glVertex3f(0.1f, -0.5f, 0.0f); glVertex3f(0.7f, -0.5f, 0.0f); glVertex3f(0.4f, 0.4f, 0.0f);
        glEnd();
    }
    glDisableClientState(GL_COLOR_ARRAY);
    glDisableClientState(GL_NORMAL_ARRAY);
    glDisableClientState(GL_VERTEX_ARRAY);
    glDeleteBuffers(1, &vbo);
    free(v);
}

/* Pixel rectangles: glRasterPos*, glDrawPixels (all four formats; alpha test, depth test against depth 0, blending,
 * depth write) and glReadPixels (RGBA / RGB, partly outside the framebuffer), mixed with geometry before and after
 * (gl_api.c:1180-1424).  What glReadPixels returns is drawn again elsewhere, so it is part of the image.
 * variant bit 0: blending on for the RGBA rectangles; bit 1: depth function GL_LEQUAL instead of GL_LESS; bit 2: the
 * raster position comes through a perspective transform (and once from behind the eye: invalid, nothing drawn) */
static void scene_pixels(int w, int h, int variant)
{
    glViewport(0, 0, w, h);
    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glOrtho(0.0, (double)w, 0.0, (double)h, -1.0, 1.0);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    glClearColor(0.15f, 0.2f, 0.1f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    glEnable(GL_DEPTH_TEST);
    glDepthFunc((variant & 2) ? GL_LEQUAL : GL_LESS);
    /* background geometry at several depths (window depth 0.25 .. 0.75; one quad exactly at depth 0) */
    glBegin(GL_QUADS);
    for (int k = 0; k < 6; k++) {
        float x0 = (float)w * 0.05f + (float)k * (float)w * 0.15f, z = (k == 3) ? 1.0f : 0.5f - 0.2f * (float)k;
        glColor4f(0.2f + 0.12f * (float)k, 0.8f - 0.1f * (float)k, 0.3f + 0.1f * (float)(k % 3), 0.6f);
        glVertex3f(x0, (float)h * 0.1f, z); glVertex3f(x0 + (float)w * 0.2f, (float)h * 0.1f, z);
        glVertex3f(x0 + (float)w * 0.2f, (float)h * 0.8f, z); glVertex3f(x0, (float)h * 0.8f, z);
    }
    glEnd();

    enum { RW = 40, RH = 24 };
    static uint8_t rgba[RW * RH * 4], rgb[RW * RH * 3], lum[RW * RH], la[RW * RH * 2];
    for (int y = 0; y < RH; y++)
        for (int x = 0; x < RW; x++) {
            int i = y * RW + x;
            rgba[i * 4] = (uint8_t)(x * 6); rgba[i * 4 + 1] = (uint8_t)(y * 10); rgba[i * 4 + 2] = (uint8_t)(255 - x * 5);
            rgba[i * 4 + 3] = (uint8_t)((x + y) * 4);
            rgb[i * 3] = (uint8_t)(255 - y * 9); rgb[i * 3 + 1] = (uint8_t)(x * 3 + y * 2); rgb[i * 3 + 2] = (uint8_t)(x ^ y) * 4;
            lum[i] = (uint8_t)((x * y) & 0xFF);
            la[i * 2] = (uint8_t)(x * 5 + 20); la[i * 2 + 1] = (uint8_t)(((x / 4 + y / 4) & 1) ? 230 : 40);
        }

    if (variant & 4) {                               /* raster position through a perspective transform */
        glMatrixMode(GL_PROJECTION);
        glLoadIdentity();
        glFrustum(-0.1 * (double)w / (double)h, 0.1 * (double)w / (double)h, -0.1, 0.1, 0.1, 100.0);
        glMatrixMode(GL_MODELVIEW);
        glRasterPos3f(0.4f, 0.2f, 3.0f);              /* behind the eye: invalid */
        glDrawPixels(RW, RH, GL_RGB, GL_UNSIGNED_BYTE, rgb);
        glRasterPos3f(-1.1f, -0.6f, -2.5f);
        glDrawPixels(RW, RH, GL_RGB, GL_UNSIGNED_BYTE, rgb);
        glMatrixMode(GL_PROJECTION);
        glLoadIdentity();
        glOrtho(0.0, (double)w, 0.0, (double)h, -1.0, 1.0);
        glMatrixMode(GL_MODELVIEW);
    }

    /* RGBA: depth test against depth 0 (passes everywhere but over the quad at depth 0 under GL_LESS), optional blend */
    if (variant & 1) { glEnable(GL_BLEND); glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA); }
    glRasterPos2i(w / 2 - 10, h / 3);
    glDrawPixels(RW, RH, GL_RGBA, GL_UNSIGNED_BYTE, rgba);
    glEnable(GL_ALPHA_TEST);
    glAlphaFunc(GL_GREATER, 0.3f);
    glRasterPos2i(w / 8, h / 2);
    glDrawPixels(RW, RH, GL_RGBA, GL_UNSIGNED_BYTE, rgba);
    glDisable(GL_BLEND);
    /* luminance + alpha, alpha test on, no depth test: partly off the left / bottom edge */
    glDisable(GL_DEPTH_TEST);
    glRasterPos2i(5, 3);
    glDrawPixels(RW, RH, GL_LUMINANCE_ALPHA, GL_UNSIGNED_BYTE, la);
    glDisable(GL_ALPHA_TEST);
    /* RGB opaque across the right / top edge */
    glRasterPos2f((float)w - 17.5f, (float)h - 10.25f);
    glDrawPixels(RW, RH, GL_RGB, GL_UNSIGNED_BYTE, rgb);
    /* luminance with depth test and depth write off, blended additively */
    glEnable(GL_DEPTH_TEST);
    glDepthMask(GL_FALSE);
    glEnable(GL_BLEND);
    glBlendFunc(GL_ONE, GL_ONE);
    glRasterPos2i(w / 2 + 40, h / 2 + 10);
    glDrawPixels(RW, RH, GL_LUMINANCE, GL_UNSIGNED_BYTE, lum);
    glDisable(GL_BLEND);
    glDepthMask(GL_TRUE);

    /* geometry after the rectangles: hidden where a rectangle wrote depth 0 */
    glBegin(GL_TRIANGLES);
    glColor3f(0.9f, 0.9f, 0.2f);
    glVertex3f((float)w * 0.1f, (float)h * 0.25f, -0.9f); glVertex3f((float)w * 0.9f, (float)h * 0.3f, -0.9f);
    glVertex3f((float)w * 0.5f, (float)h * 0.75f, -0.9f);
    glEnd();

    /* read two rectangles back (one hanging over the left / top edges) and draw them again */
    static uint8_t back4[48 * 32 * 4], back3[48 * 32 * 3];
    glReadPixels(w / 2 - 24, h / 3 - 4, 48, 32, GL_RGBA, GL_UNSIGNED_BYTE, back4);
    glReadPixels(-9, h - 20, 48, 32, GL_RGB, GL_UNSIGNED_BYTE, back3);
    glDisable(GL_DEPTH_TEST);
    glRasterPos2i(w - 60, 8);
    glDrawPixels(48, 32, GL_RGBA, GL_UNSIGNED_BYTE, back4);
    glRasterPos2i(w / 3, h - 40);
    glDrawPixels(48, 32, GL_RGB, GL_UNSIGNED_BYTE, back3);
    glEnable(GL_DEPTH_TEST);
    glBegin(GL_TRIANGLES);                            /* and something on top of it all */
    glColor3f(0.2f, 0.3f, 0.95f);
    glVertex3f((float)w * 0.6f, 4.0f, -0.95f); glVertex3f((float)w - 4.0f, 6.0f, -0.95f); glVertex3f((float)w * 0.8f, (float)h * 0.3f, -0.95f);
    glEnd();
}

/* Display lists whose glBegin ... glEnd stretches are long enough to be compiled into array draws (the front end's
 * ListRun path): the result must be what the reference's call-by-call replay gives.
 * variant 0: lit Suzanne (normals set before every vertex) called four times under different transforms, with the
 *            current colour / material changed between calls; the list also holds state changes around the geometry
 *         1: a strip whose colour is set inside the run before the first vertex, texture coordinates never (they come
 *            from the caller), and a second run whose colour is first set AFTER its first vertex (must be replayed)
 *         2: GL_COLOR_MATERIAL with per-vertex colours inside the run (left to the per-call path), two-sided lighting
 *         3: a list redefined after it was drawn from, nested lists, glCallLists with a list base, COMPILE_AND_EXECUTE */
static void scene_displaylist_runs(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.05f, 0.05f, 0.1f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    GLuint base = glGenLists(4);
    if (variant == 0 || variant == 3) {
        suzanne_lights_and_material();
        glNewList(base, variant == 3 ? GL_COMPILE_AND_EXECUTE : GL_COMPILE);
        glPushMatrix();
        glRotatef(25.0f, 0.0f, 1.0f, 0.0f);
        suzanne_immediate();
        glPopMatrix();
        glShadeModel(GL_FLAT);
        glBegin(GL_QUADS);                      /* a short run: replayed */
        glNormal3f(0.0f, 0.0f, 1.0f);
        glVertex3f(-1.2f, -1.1f, 0.2f); glVertex3f(1.2f, -1.1f, 0.2f); glVertex3f(1.2f, -0.9f, 0.2f); glVertex3f(-1.2f, -0.9f, 0.2f);
        glEnd();
        glShadeModel(GL_SMOOTH);
        glEndList();
        for (int k = 0; k < 4; k++) {
            GLfloat md[4] = { 0.3f + 0.2f * (float)k, 0.8f - 0.15f * (float)k, 0.4f, 1.0f };
            glMaterialfv(GL_FRONT, GL_DIFFUSE, md);
            glLoadIdentity();
            glTranslatef(-2.7f + 1.8f * (float)k, (k & 1) ? 0.5f : -0.4f, -6.0f - 0.5f * (float)k);
            if (variant == 3 && k == 2) {       /* redefine the list the queued draws were made from */
                glNewList(base, GL_COMPILE);
                glScalef(0.6f, 1.3f, 0.6f);
                suzanne_immediate();
                glEndList();
            }
            if (variant == 3 && k == 3) {
                glNewList(base + 1, GL_COMPILE);
                glCallList(base);
                glTranslatef(0.0f, -1.6f, 0.0f);
                glCallList(base);
                glEndList();
                GLubyte which[2] = { 1, 0 };
                glListBase(base);
                glCallLists(2, GL_UNSIGNED_BYTE, which);
                glListBase(0);
            } else glCallList(base);
        }
    } else if (variant == 1) {
        GLuint tex = make_checker_rgb(32, 4, 0);
        (void)tex;
        glEnable(GL_TEXTURE_2D);
        glNewList(base, GL_COMPILE);
        glBegin(GL_TRIANGLE_STRIP);
        glColor4f(0.9f, 0.7f, 0.3f, 1.0f);      /* before the first vertex: compiled */
        for (int k = 0; k < 40; k++) {
            float a = 0.16f * (float)k;
            if (k == 20) glColor4f(0.3f, 0.6f, 0.95f, 1.0f);
            glVertex3f(-2.5f + 0.125f * (float)k, 0.9f + 0.5f * sinf(a), 0.0f);
            glVertex3f(-2.5f + 0.125f * (float)k, 0.2f + 0.3f * cosf(a), 0.3f);
        }
        glEnd();
        glBegin(GL_TRIANGLE_FAN);
        glVertex3f(0.0f, -1.0f, 0.0f);          /* this vertex uses the caller's colour: the run is replayed */
        for (int k = 0; k <= 30; k++) {
            glColor3f(0.03f * (float)k, 1.0f - 0.03f * (float)k, 0.5f);
            glTexCoord2f(0.1f * (float)k, 0.5f);
            glVertex3f(1.4f * cosf(0.2094f * (float)k), -1.0f + 0.8f * sinf(0.2094f * (float)k), 0.0f);
        }
        glEnd();
        glEndList();
        for (int k = 0; k < 3; k++) {
            glLoadIdentity();
            glTranslatef(-1.0f + 1.0f * (float)k, 0.2f * (float)k, -5.0f - (float)k);
            glTexCoord2f(0.25f * (float)k, 0.6f);               /* the compiled strip takes these */
            glColor3f(1.0f, 0.2f * (float)k, 1.0f);
            glCallList(base);
        }
        glDisable(GL_TEXTURE_2D);
    } else {
        suzanne_lights_and_material();
        glLightModeli(GL_LIGHT_MODEL_TWO_SIDE, 1);
        glEnable(GL_COLOR_MATERIAL);
        glColorMaterial(GL_FRONT_AND_BACK, GL_DIFFUSE);
        glNewList(base, GL_COMPILE);
        glBegin(GL_TRIANGLES);
        for (int i = 0; i < g_mesh_nf && i < 400; i++)
            for (int j = 0; j < 3; j++) {
                int vi = g_mesh_faces[i * 3 + j];
                glColor3f(0.5f + 0.5f * g_mesh_nrm[vi * 3], 0.5f + 0.5f * g_mesh_nrm[vi * 3 + 1], 0.6f);
                glNormal3f(g_mesh_nrm[vi * 3], g_mesh_nrm[vi * 3 + 1], g_mesh_nrm[vi * 3 + 2]);
                glVertex3f(g_mesh_pos[vi * 3], g_mesh_pos[vi * 3 + 1], g_mesh_pos[vi * 3 + 2]);
            }
        glEnd();
        glEndList();
        for (int k = 0; k < 2; k++) {
            glLoadIdentity();
            glTranslatef(-1.2f + 2.4f * (float)k, 0.0f, -4.5f);
            glRotatef(160.0f * (float)k, 0.0f, 1.0f, 0.0f);
            glCallList(base);
        }
        glDisable(GL_COLOR_MATERIAL);
        glBegin(GL_TRIANGLES);                   /* the material the list left behind */
        glNormal3f(0.0f, 0.0f, 1.0f);
        glVertex3f(-0.5f, -1.6f, -4.0f); glVertex3f(0.5f, -1.6f, -4.0f); glVertex3f(0.0f, -1.0f, -4.0f);
        glEnd();
    }
    glDeleteLists(base, 4);
}

/* Suzanne as an INDEXED mesh -- 507 vertices (position + normal, texture coordinates derived from
 * the position) in one buffer, 968 x 3 indices drawn with glDrawElements (gl_api.c:1854-1941): every vertex is referenced
 * 5.7 times.  variant 0: GL_UNSIGNED_SHORT indices in an element buffer; 1: GL_UNSIGNED_INT in an element buffer;
 * 2: GL_UNSIGNED_SHORT indices in client memory; 3: GL_UNSIGNED_INT, three draws of the mesh under different modelview
 * matrices, two-sided lighting; 4: a 5x5 grid of vertices with ubyte colours under GL_COLOR_MATERIAL drawn as indexed
 * GL_TRIANGLE_STRIPs with GL_UNSIGNED_BYTE indices; 5: variant 0 with GL_PHONG shading (per-fragment lighting);
 * 6: GL_UNSIGNED_SHORT, textured + fog, polygon mode GL_LINE for back faces. */
static void scene_indexed(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 100.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.15f, 0.15f, 0.2f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    GLuint buf[2];
    glGenBuffers(2, buf);
    if (variant == 4) {
        struct { float x, y, z; uint8_t c[4]; } grid[25];
        for (int j = 0; j < 5; j++)
            for (int i = 0; i < 5; i++) {
                int k = j * 5 + i;
                grid[k].x = (float)i * 0.5f - 1.0f; grid[k].y = (float)j * 0.5f - 1.0f; grid[k].z = 0.15f * (float)((i * 7 + j * 3) % 5);
                grid[k].c[0] = (uint8_t)(40 + 50 * i); grid[k].c[1] = (uint8_t)(30 + 45 * j); grid[k].c[2] = (uint8_t)(255 - 30 * i - 20 * j); grid[k].c[3] = 255;
            }
        uint8_t strip[4][10];
        for (int j = 0; j < 4; j++)
            for (int i = 0; i < 5; i++) { strip[j][2 * i] = (uint8_t)(j * 5 + i); strip[j][2 * i + 1] = (uint8_t)((j + 1) * 5 + i); }
        glBindBuffer(GL_ARRAY_BUFFER, buf[0]);
        glBufferData(GL_ARRAY_BUFFER, sizeof grid, grid, GL_STATIC_DRAW);
        glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, buf[1]);
        glBufferData(GL_ELEMENT_ARRAY_BUFFER, sizeof strip, strip, GL_STATIC_DRAW);
        glEnableClientState(GL_VERTEX_ARRAY);
        glEnableClientState(GL_COLOR_ARRAY);
        glVertexPointer(3, GL_FLOAT, 16, (const void *)0);
        glColorPointer(4, GL_UNSIGNED_BYTE, 16, (const void *)12);
        suzanne_lights_and_material();
        glEnable(GL_COLOR_MATERIAL);
        glColorMaterial(GL_FRONT_AND_BACK, GL_AMBIENT_AND_DIFFUSE);
        glNormal3f(0.0f, 0.0f, 1.0f);
        glLoadIdentity();
        glTranslatef(0.0f, 0.0f, -3.2f);
        glRotatef(-25.0f, 1.0f, 0.0f, 0.0f);
        for (int j = 0; j < 4; j++) glDrawElements(GL_TRIANGLE_STRIP, 10, GL_UNSIGNED_BYTE, (const void *)(size_t)(j * 10));
        glDisableClientState(GL_COLOR_ARRAY);
        glDisableClientState(GL_VERTEX_ARRAY);
        glDeleteBuffers(2, buf);
        return;
    }
    /* 507 vertices, interleaved position3 + normal3 + uv2 */
    float *vb = (float *)malloc((size_t)g_mesh_nv * 8 * sizeof(float));
    for (int i = 0; i < g_mesh_nv; i++) {
        memcpy(vb + i * 8, g_mesh_pos + i * 3, 12);
        memcpy(vb + i * 8 + 3, g_mesh_nrm + i * 3, 12);
        vb[i * 8 + 6] = (g_mesh_pos[i * 3] + 0.7f) * 4.0f; vb[i * 8 + 7] = (g_mesh_pos[i * 3 + 1] + 0.5f) * 4.0f;
    }
    int ni = g_mesh_nf * 3;
    uint16_t *i16 = (uint16_t *)malloc((size_t)ni * 2);
    uint32_t *i32 = (uint32_t *)malloc((size_t)ni * 4);
    for (int i = 0; i < ni; i++) { i16[i] = (uint16_t)g_mesh_faces[i]; i32[i] = (uint32_t)g_mesh_faces[i]; }
    glBindBuffer(GL_ARRAY_BUFFER, buf[0]);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)((size_t)g_mesh_nv * 32), vb, GL_STATIC_DRAW);
    const int wide = (variant == 1 || variant == 3);
    const void *indices = (const void *)0;
    if (variant == 2) indices = i16;
    else {
        glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, buf[1]);
        if (wide) glBufferData(GL_ELEMENT_ARRAY_BUFFER, (GLsizeiptr)((size_t)ni * 4), i32, GL_STATIC_DRAW);
        else glBufferData(GL_ELEMENT_ARRAY_BUFFER, (GLsizeiptr)((size_t)ni * 2), i16, GL_STATIC_DRAW);
    }
    glEnableClientState(GL_VERTEX_ARRAY);
    glEnableClientState(GL_NORMAL_ARRAY);
    glVertexPointer(3, GL_FLOAT, 32, (const void *)0);
    glNormalPointer(GL_FLOAT, 32, (const void *)12);
    suzanne_lights_and_material();
    if (variant == 5) glShadeModel(GL_PHONG);
    if (variant == 3) {
        GLfloat bd[] = { 0.1f, 0.6f, 0.9f, 1.0f };
        glMaterialfv(GL_BACK, GL_DIFFUSE, bd);
        glLightModeli(GL_LIGHT_MODEL_TWO_SIDE, 1);
    }
    if (variant == 6) {
        glEnableClientState(GL_TEXTURE_COORD_ARRAY);
        glTexCoordPointer(2, GL_FLOAT, 32, (const void *)24);
        glEnable(GL_TEXTURE_2D);
        make_checker_rgb(64, 4, 1);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_REPEAT);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_REPEAT);
        GLfloat fog[] = { 0.15f, 0.15f, 0.2f, 1.0f };
        glEnable(GL_FOG); glFogi(GL_FOG_MODE, GL_EXP2); glFogf(GL_FOG_DENSITY, 0.25f); glFogfv(GL_FOG_COLOR, fog);
        glPolygonMode(GL_BACK, GL_LINE);
    }
    const int draws = (variant == 3) ? 3 : 1;
    for (int k = 0; k < draws; k++) {
        glLoadIdentity();
        glTranslatef(draws == 1 ? 0.0f : (float)(k - 1) * 1.6f, 0.0f, draws == 1 ? -2.2f : -4.0f);
        glRotatef(20.0f + 70.0f * (float)k, 1.0f, 0.0f, 0.0f);
        glRotatef(30.0f, 0.0f, 1.0f, 0.0f);
        glDrawElements(GL_TRIANGLES, ni, wide ? GL_UNSIGNED_INT : GL_UNSIGNED_SHORT, indices);
    }
    i16[0] = 1; /* client-memory indices are consumed at the call: must not affect the draw above */
    glDisableClientState(GL_TEXTURE_COORD_ARRAY);
    glDisableClientState(GL_NORMAL_ARRAY);
    glDisableClientState(GL_VERTEX_ARRAY);
    glDeleteBuffers(2, buf);
    free(vb); free(i16); free(i32);
}

/* Object lifetime ("observationally synchronous"): the reference executes every draw at the call, so an
 * object may be deleted, redefined or have its name reused right after the draw that used it.  A back end that queues
 * draws must behave the same.  variant 0: glDeleteTextures / glDeleteBuffers / glDeleteLists directly behind the draws
 * that use them, then the names are generated again, filled with different contents and drawn; 1: glTexImage2D /
 * glBufferData / glBufferSubData / glNewList on objects that queued draws still use; 2: as 0 inside a display list that
 * is deleted while it ... has just been called. */
static void scene_lifetime(int w, int h, int variant)
{
    frustum_like_testbed(w, h, 50.0);
    glEnable(GL_DEPTH_TEST);
    glClearColor(0.05f, 0.1f, 0.1f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
    static const float quad[4][5] = { { -1, -1, 0, 0, 0 }, { 1, -1, 0, 2, 0 }, { 1, 1, 0, 2, 2 }, { -1, 1, 0, 0, 2 } };
    static const float small[4][5] = { { -0.5f, -0.5f, 0, 0, 0 }, { 0.5f, -0.5f, 0, 1, 0 }, { 0.5f, 0.5f, 0, 1, 1 }, { -0.5f, 0.5f, 0, 0, 1 } };
    glEnableClientState(GL_VERTEX_ARRAY);
    glEnableClientState(GL_TEXTURE_COORD_ARRAY);
    glEnable(GL_TEXTURE_2D);
    glTexEnvi(GL_TEXTURE_ENV, GL_TEXTURE_ENV_MODE, GL_REPLACE);
    for (int round = 0; round < 3; round++) {
        GLuint tex = round & 1 ? make_radial_rgba(32) : make_checker_rgb(64, 4 << round, round & 1);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, round == 2 ? GL_NEAREST : GL_LINEAR);
        GLuint vbo;
        glGenBuffers(1, &vbo);
        glBindBuffer(GL_ARRAY_BUFFER, vbo);
        glBufferData(GL_ARRAY_BUFFER, sizeof quad, quad, GL_STATIC_DRAW);
        glVertexPointer(3, GL_FLOAT, 20, (const void *)0);
        glTexCoordPointer(2, GL_FLOAT, 20, (const void *)12);
        glLoadIdentity();
        glTranslatef((float)(round - 1) * 1.7f, 0.4f, -5.0f - (float)round * 0.3f);
        glRotatef(15.0f * (float)round, 0.0f, 0.0f, 1.0f);
        glDrawArrays(GL_QUADS, 0, 4);
        GLuint list = glGenLists(1);
        glNewList(list, GL_COMPILE);
        glBegin(GL_TRIANGLES);
        for (int k = 0; k < 30; k++) {      /* long enough to become a compiled run */
            float a = (float)k * 0.2094395f;
            glTexCoord2f(0.5f, 0.5f); glVertex3f(0.0f, -1.6f, 0.0f);
            glTexCoord2f(0.5f + 0.5f * cosf(a), 0.5f + 0.5f * sinf(a)); glVertex3f(0.6f * cosf(a), -1.6f + 0.6f * sinf(a), 0.0f);
            glTexCoord2f(0.5f + 0.5f * cosf(a + 0.2094395f), 0.5f + 0.5f * sinf(a + 0.2094395f));
            glVertex3f(0.6f * cosf(a + 0.2094395f), -1.6f + 0.6f * sinf(a + 0.2094395f), 0.0f);
        }
        glEnd();
        glEndList();
        glCallList(list);
        if (variant == 1) {
            /* redefinition while the draws above are still queued */
            uint8_t px[16 * 16 * 3];
            for (int i = 0; i < 16 * 16; i++) { px[i * 3] = (uint8_t)(i * 7); px[i * 3 + 1] = (uint8_t)(255 - i); px[i * 3 + 2] = (uint8_t)(round * 90); }
            glTexImage2D(GL_TEXTURE_2D, 0, GL_RGB, 16, 16, 0, GL_RGB, GL_UNSIGNED_BYTE, px);
            glBufferData(GL_ARRAY_BUFFER, sizeof small, small, GL_STATIC_DRAW);
            glLoadIdentity();
            glTranslatef((float)(round - 1) * 1.7f, 0.4f, -4.0f);
            glDrawArrays(GL_QUADS, 0, 4);
            float moved[5] = { -0.9f, -0.9f, 0.2f, 0, 0 };
            glBufferSubData(GL_ARRAY_BUFFER, 0, sizeof moved, moved);
            glTranslatef(0.0f, 1.3f, 0.0f);
            glDrawArrays(GL_QUADS, 0, 4);
            glNewList(list, GL_COMPILE);
            glBegin(GL_QUADS);
            glTexCoord2f(0, 0); glVertex3f(-0.3f, -2.0f, 0.5f); glTexCoord2f(1, 0); glVertex3f(0.3f, -2.0f, 0.5f);
            glTexCoord2f(1, 1); glVertex3f(0.3f, -1.4f, 0.5f); glTexCoord2f(0, 1); glVertex3f(-0.3f, -1.4f, 0.5f);
            glEnd();
            glEndList();
            glCallList(list);
        }
        /* the objects go away while everything above may still be queued; their names are reused by the next round */
        glDeleteTextures(1, &tex);
        glDeleteBuffers(1, &vbo);
        glDeleteLists(list, 1);
        if (variant == 2 && round == 1) {    /* drawing with the deleted names bound: no texture, no buffer -> nothing / untextured */
            glBindTexture(GL_TEXTURE_2D, tex);
            glBegin(GL_TRIANGLES);
            glColor3f(0.9f, 0.3f, 0.1f);
            glVertex3f(-0.4f, 1.2f, 0.3f); glVertex3f(0.4f, 1.2f, 0.3f); glVertex3f(0.0f, 1.8f, 0.3f);
            glEnd();
            glCallList(list);
            glColor3f(1.0f, 1.0f, 1.0f);
        }
    }
    glDisableClientState(GL_TEXTURE_COORD_ARRAY);
    glDisableClientState(GL_VERTEX_ARRAY);
}

/* ---------------------------------------------------------------- registry */
typedef void (*scene_fn)(int, int, int);
static const struct { const char *name; scene_fn fn; } g_scenes[] = {
    { "c1_suzanne", scene_c1 },
    { "c2_cube", scene_c2_cube },
    { "c2_floor", scene_c2_floor },
    { "c2_texenv", scene_c2_texenv },
    { "c3_fill", scene_c3 },
    { "c4_grid", scene_c4 },
    { "clipping", scene_clipping },
    { "primitives", scene_primitives },
    { "zbuffer", scene_zbuffer },
    { "fog", scene_fog },
    { "blend", scene_blend },
    { "stencil", scene_stencil },
    { "lighting", scene_lighting },
    { "displaylist", scene_displaylist },
    { "vbo", scene_vbo },
    { "validation", scene_validation },
    { "scissor", scene_scissor },
    { "texture_misc", scene_texture_misc },
    { "state_churn", scene_state_churn },
    { "lines", scene_lines },
    { "wireframe", scene_wireframe },
    { "depth_order", scene_depth_order },
    { "cull", scene_cull },
    { "vbo_large", scene_vbo_large },
    { "pixels", scene_pixels },
    { "displaylist_runs", scene_displaylist_runs },
    { "indexed", scene_indexed },
    { "lifetime", scene_lifetime },
};

int scene_count(void) { return (int)(sizeof g_scenes / sizeof g_scenes[0]); }
const char *scene_name(int i) { return (i >= 0 && i < scene_count()) ? g_scenes[i].name : NULL; }

/* Issue every GL call of the named scene into the current context.  Returns 0, or -1 for an unknown name. */
int scene_render(const char *name, int w, int h, int variant)
{
    for (int i = 0; i < scene_count(); i++)
        if (strcmp(g_scenes[i].name, name) == 0) {
            g_scenes[i].fn(w, h, variant);
            return 0;
        }
    return -1;
}
