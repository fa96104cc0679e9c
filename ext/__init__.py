"""Drop-in `ext` namespace of the reference (ext.lab2im, ext.neuron) backed by the B200 engine (synthsr_b200)."""
