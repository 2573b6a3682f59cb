"""B200 re-host of reference scripts/extract_retrieval_features.py -- same flags; see freepose_b200/cli.py."""
from freepose_b200.cli import run_extract_retrieval_features

if __name__ == "__main__":
    run_extract_retrieval_features()
