// Explicit instantiation of the witness engine for Pallas (see engine.cuh).
#include "engine.cuh"
namespace eagen {
IEngine* make_engine_pallas(int device) { return new Engine<Pallas>(device); }
}  // namespace eagen
