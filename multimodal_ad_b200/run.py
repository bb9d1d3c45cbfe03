"""Launcher that runs one of the reference's scripts UNCHANGED on the accelerated modules.

    cd /path/to/Multimodal_AD && python -m multimodal_ad_b200.run train_ResNet3D.py [args...]

`python script.py` puts the script's own directory first on sys.path, so the reference's `models/` package would always
win over PYTHONPATH; this launcher puts the shim package (multimodal_ad_b200/shim, see its models/__init__.py) in front,
then executes the script as __main__ exactly as `python script.py` would.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m multimodal_ad_b200.run <script.py> [args...]")
    script = os.path.abspath(argv[0])
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
    sys.path[:0] = [shim, os.path.dirname(script)]         # shim first, then what `python script.py` would have put first
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
