// Explicit instantiation of the witness engine for Grumpkin (see engine.cuh).
#include "engine.cuh"
namespace eagen {
#ifdef EAGEN_DEV_PALLAS_ONLY   // development builds for A/B timing only (tools/variant.sh): never shipped
IEngine* make_engine_grumpkin(int) { throw StatusError{EAGEN_E_ARG, "this development build holds the Pallas engine only"}; }
#else
IEngine* make_engine_grumpkin(int device) { return new Engine<Grumpkin>(device); }
#endif
}  // namespace eagen
