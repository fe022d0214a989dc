// Empty stand-in for OWL's host API header (extern/owl/owl/include/owl/owl.h).
// The oracle build only needs OWL's header-only vector math, which is included
// from the reference tree itself; the host API (OptiX context management) is not
// part of the per-path code being restated.
#pragma once
