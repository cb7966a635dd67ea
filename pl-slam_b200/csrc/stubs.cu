// Temporary no-op destructors for subsystems that are not built yet (removed as they land).
#include "plf_internal.h"
extern "C" {
void plf_pipe_free(plf_ctx*) {}
}
