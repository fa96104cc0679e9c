"""Extracts (by AST, no import) the public signatures of the reference's boundary functions into
tests/golden/reference_signatures.json.  Build container only (needs /root/reference)."""
import ast
import json
import os

REF = '/root/reference'
TARGETS = {'SynthSR/training.py': ['training'], 'SynthSR/brain_generator.py': ['BrainGenerator.__init__'],
           'SynthSR/labels_to_image_model.py': ['labels_to_image_model', 'get_shapes'],
           'SynthSR/model_inputs.py': ['build_model_inputs'], 'SynthSR/metrics_model.py': ['metrics_model'],
           'ext/neuron/models.py': ['unet'],
           'SynthSR/fine_tuning_with_adversary.py': ['training', 'make_discriminator'],
           'ext/lab2im/utils.py': ['load_volume', 'save_volume', 'get_volume_info', 'get_list_labels', 'reformat_to_list',
                                   'get_padding_margin', 'build_training_generator', 'draw_value_from_distribution']}


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return [[n, d] for n, d in zip(names, defaults) if n != 'self']


out = {}
for path, names in TARGETS.items():
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for name in names:
        if '.' in name:
            cls, meth = name.split('.')
            c = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0]
            fn = [n for n in c.body if isinstance(n, ast.FunctionDef) and n.name == meth][0]
        else:
            fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
        out['%s:%s' % (path, name)] = sig(fn)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_signatures.json'), 'w'), indent=1)
print({k: len(v) for k, v in out.items()})
