// Explicit instantiation of the witness engine for Vesta (see engine.cuh).
#include "engine.cuh"
namespace eagen {
IEngine* make_engine_vesta(int device) { return new Engine<Vesta>(device); }
}  // namespace eagen
