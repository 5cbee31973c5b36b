// TEST INFRASTRUCTURE — the reference's own voice-activity test, compiled from where it lies.
//
// _high_pass_filter / _vad_simple live in /root/reference/src/speech_to_text.cpp:53-104, in a translation unit that needs the Godot
// engine to run.  oracle/Makefile cuts exactly those two function definitions out of that file into _ref/vad_ref_extract.inc (a build
// product, git-ignored: no reference source enters the repo) and this file compiles them against a minimal stand-in for the three
// godot-cpp names they use.  The arithmetic is the reference's, statement for statement.
#include <cmath>
#include <cstddef>
#include <vector>

#define Math_PI 3.1415926535897932384626433833     /* thirdparty/godot-cpp/include/godot_cpp/core/math.hpp:45 */

namespace {
struct PackedFloat32Array {                         // the accessors the two functions use
    float * p; size_t n;
    size_t size() const { return n; }
    float & operator[](size_t i) { return p[i]; }
};
struct UtilityFunctions { template <class... A> static void print(A...) {} };
inline int rtos(double) { return 0; }

#include "_ref/vad_ref_extract.inc"
}  // namespace

extern "C" __attribute__((visibility("default"))) void ref_high_pass_filter(float * data, int n, float cutoff, float sample_rate) {
    PackedFloat32Array a{data, (size_t) n};
    _high_pass_filter(a, cutoff, sample_rate);
}
extern "C" __attribute__((visibility("default"))) int ref_vad_simple(float * pcm, int n, int sample_rate, int last_ms, float vad_thold, float freq_thold) {
    PackedFloat32Array a{pcm, (size_t) n};
    return _vad_simple(a, sample_rate, last_ms, vad_thold, freq_thold, false) ? 1 : 0;
}
