"""B200 re-host of reference scripts/dino_inference.py -- same flags; see freepose_b200/cli.py."""
from freepose_b200.cli import run_dino_inference

if __name__ == "__main__":
    run_dino_inference()
