"""TEST-ONLY stand-in for `pooch` (absent from the image, no network): the reference's example generators fetch their CSVs
from github (gempy/API/initialization_API.py:215-243); the same files ship with the reference under examples/data, so a
URL is resolved to that local copy."""
import hashlib
import os

REFERENCE = os.environ.get("GEMPY_REFERENCE", "/root/reference")
URL_PREFIX = "https://raw.githubusercontent.com/cgre-aachen/gempy_data/master/"


def retrieve(url, known_hash=None, **kwargs):
    if not url.startswith(URL_PREFIX):
        raise OSError(f"no network: cannot fetch {url}")
    path = os.path.join(REFERENCE, "examples", url[len(URL_PREFIX):].lstrip("/"))
    if not os.path.exists(path):
        raise OSError(f"{url}: no local copy at {path}")
    return path


def file_hash(path, alg="sha256"):
    h = hashlib.new(alg)
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()
