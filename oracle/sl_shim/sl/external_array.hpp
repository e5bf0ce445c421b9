// Minimal stand-in for <sl/external_array.hpp>: the reference includes it but uses
// nothing from it on the svbuilder path (see cstdint.hpp header note).
#pragma once
#include <sl/cstdint.hpp>
#include <sl/clock.hpp>
#include <sl/axis_aligned_box.hpp>
