"""SynthSR-Hyperfine prediction (T1 + T2 at 1.5 x 1.5 x 5 mm -> 1 mm MP-RAGE): same arguments as the reference's
scripts/predict_command_line_hyperfine.py (`--cpu` / `--threads` configure TensorFlow there and are not supported here).

    python scripts/predict_command_line_hyperfine.py <T1 image|folder> <T2 image|folder> <prediction|folder>
"""
import os
import sys
from argparse import ArgumentParser

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == '__main__':
    print('\n')
    print('SynthSR-Hyperfine prediction')
    print('\n')
    parser = ArgumentParser()
    parser.add_argument("path_t1_images", type=str, help="T1 images to super-resolve, at native 1.5x1.5x5 axial resolution.")
    parser.add_argument("path_t2_images", type=str, help="T2 images (registered to the T1s; they are resampled onto them).")
    parser.add_argument("path_predictions", type=str, help="path where to save the synthetic 1mm MP-RAGEs.")
    parser.add_argument("--cpu", action="store_true", help="not supported by this engine (CUDA only).")
    parser.add_argument("--threads", type=int, default=1, dest="threads", help="ignored (TensorFlow CPU setting).")
    parser.add_argument("--model", default=None, help="(optional) Use a different model file.")
    args = parser.parse_args()
    if args.cpu:
        raise SystemExit('this engine runs on CUDA devices only (no CPU fallback)')
    from SynthSR.predict import predict_hyperfine
    predict_hyperfine(args.path_t1_images, args.path_t2_images, args.path_predictions, model=args.model)
    print(' ')
    print('All done!')
    print(' ')
