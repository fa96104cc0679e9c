"""SynthSR prediction from the terminal: same arguments as the reference's scripts/predict_command_line.py (its `--cpu` /
`--threads` flags configure TensorFlow and have no meaning for this engine: there is no CPU fallback).

    python scripts/predict_command_line.py <image or folder> <prediction or folder> [--ct] [--model x.h5] [--disable_flipping]
"""
import os
import sys
from argparse import ArgumentParser

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == '__main__':
    print('\n')
    print('SynthSR prediction')
    print('\n')
    parser = ArgumentParser()
    parser.add_argument("path_images", type=str,
                        help="images to super-resolve / synthesize. Can be the path to a single image or to a folder")
    parser.add_argument("path_predictions", type=str,
                        help="path where to save the synthetic 1mm MP-RAGEs. Must be the same type "
                             "as path_images (path to a single image or to a folder)")
    parser.add_argument("--cpu", action="store_true", help="not supported by this engine (CUDA only).")
    parser.add_argument("--threads", type=int, default=1, dest="threads", help="ignored (TensorFlow CPU setting).")
    parser.add_argument("--ct", action="store_true", help="use this flag for ct scans.")
    parser.add_argument("--model", default=None, help="(optional) Use a different model file.")
    parser.add_argument("--disable_flipping", action="store_true",
                        help="(optional) Use this flag to disable flipping augmentation at test time.")
    args = parser.parse_args()
    if args.cpu:
        raise SystemExit('this engine runs on CUDA devices only (no CPU fallback)')
    from SynthSR.predict import predict
    predict(args.path_images, args.path_predictions, model=args.model, ct=args.ct, disable_flipping=args.disable_flipping)
    print(' ')
