// render_hair_msnn <config.json> [BETA] — headless stand-in for the reference executable of the
// same name (render_hair_msnn.cu: main()).  See hm_main_common.h.
#include "hm_main_common.h"
int main(int argc, char** argv) { return hm_main(argc, argv, HM_RENDER_HAIR_MSNN, "render_hair_msnn"); }
