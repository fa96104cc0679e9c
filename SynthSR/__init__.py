"""Drop-in `SynthSR` package of the reference (training / brain_generator API) backed by the B200 engine."""
