// render_path_tracing <config.json> [BETA] — headless stand-in for the reference executable of the
// same name (render_path_tracing.cu: main()).  See hm_main_common.h.
#include "hm_main_common.h"
int main(int argc, char** argv) { return hm_main(argc, argv, HM_RENDER_PATH_TRACING, "render_path_tracing"); }
