"""CPU test: the shipped replacement for the reference's SDL glue (include/mytinygl/sdl.h) compiles as C99 against
a minimal SDL2 stand-in, defines the reference's entry points, and a testbed-style main() links against the product library."""
import subprocess

from mytinygl_b200 import REPO_ROOT

PROGRAM = r"""
#include <mytinygl/sdl.h>
int main(void)
{
    if (mtgl_init("testbed", 320, 240) != 0) return 1;
    glClearColor(0.1f, 0.2f, 0.3f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT);
    mtgl_swap();
    mtgl_destroy();
    return mtgl_window != NULL || mtgl_renderer != NULL || mtgl_texture != NULL || mtgl_ctx != NULL;
}
"""


def test_sdl_header_compiles_and_binds(tmp_path):
    src = tmp_path / "main.c"
    src.write_text(PROGRAM)
    obj = tmp_path / "main.o"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-c", str(src), "-o", str(obj),
           f"-I{REPO_ROOT / 'include'}", f"-I{REPO_ROOT / 'tests' / 'stubs'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms = subprocess.run(["nm", str(obj)], capture_output=True, text=True).stdout
    for need in ("gl_create_context", "gl_make_current", "gl_destroy_context", "mtgl_map_framebuffer", "SDL_UpdateTexture", "SDL_RenderPresent"):
        assert f" U {need}" in syms, need
