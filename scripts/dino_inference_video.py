"""B200 re-host of reference scripts/dino_inference_video.py -- same flags; see freepose_b200/cli.py."""
from freepose_b200.cli import run_dino_inference_video

if __name__ == "__main__":
    run_dino_inference_video()
