/* oracle/ref_prelude.h — TEST INFRASTRUCTURE (not product code).
 *
 * Force-included (-include) in front of every reference translation unit when
 * building oracle/_ref/libisx_ref.so.  The reference's Renderer.h
 * (/root/reference/cpp/Renderer.h:8-10) refuses to compile off Windows; the
 * simulation core never calls the renderer on the step path.  We pull the
 * standard headers Renderer.h needs FIRST (so they never see the macro), then
 * include Renderer.h once with _WIN32 defined (it only declares a class; no
 * Windows headers), then remove the macro again.  `#pragma once` in Renderer.h
 * makes every later `#include "Renderer.h"` a no-op.  No reference source text
 * is copied or edited.
 */
#pragma once
#include <vector>
#include <memory>
#include <string>
#include <array>
#include <cmath>
#include <utility>
#include <unordered_map>
#include <algorithm>
#include <limits>
#include <random>
#include <cstdint>
#ifndef _WIN32
#  define _WIN32 1
#  define ISX_UNDEF_WIN32 1
#endif
#include "Renderer.h"
#ifdef ISX_UNDEF_WIN32
#  undef _WIN32
#  undef ISX_UNDEF_WIN32
#endif
