// Explicit instantiation of the witness engine for Grumpkin (see engine.cuh).
#include "engine.cuh"
namespace eagen {
IEngine* make_engine_grumpkin(int device) { return new Engine<Grumpkin>(device); }
}  // namespace eagen
