import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_rust():
    return json.load(open(os.path.join(GOLDEN, "rust_tests.json")))


@pytest.fixture(scope="session")
def golden_sigs():
    return json.load(gzip.open(os.path.join(GOLDEN, "sigs.json.gz"), "rt"))


@pytest.fixture(scope="session")
def golden_kmers():
    return json.load(gzip.open(os.path.join(GOLDEN, "kmers.json.gz"), "rt"))


@pytest.fixture(scope="session")
def golden_search():
    return json.load(open(os.path.join(GOLDEN, "search.json")))


def fasta_path(name):
    return os.path.join(GOLDEN, "fasta", name)
