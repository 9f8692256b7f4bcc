#pragma once
// placeholder: ImageHelpers.cpp is not compiled into oracle/_ref (its two entry points that the
// renderer TU needs are provided by ref_harness.cpp from pre-converted raw tables)
