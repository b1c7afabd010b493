/* oracle/ref_render_stub.cpp — TEST INFRASTRUCTURE (not product code).
 * Headless definitions of the four window methods IntersectionEnv.h:82-87
 * declares (their real bodies live in the Windows-only
 * IntersectionEnv_render.cpp / Renderer.cpp, which are not on the step path). */
#include "IntersectionEnv.h"
Renderer::Renderer() {}
Renderer::~Renderer() {}
struct Renderer::Impl {};
void IntersectionEnv::render(bool, bool) {}
bool IntersectionEnv::window_should_close() const { return true; }
void IntersectionEnv::poll_events() const {}
bool IntersectionEnv::key_pressed(int) const { return false; }
