// Empty stand-in for <optix_device.h>; the intrinsics the reference names are
// provided by oracle/ref_optix_emul.h.
#pragma once
