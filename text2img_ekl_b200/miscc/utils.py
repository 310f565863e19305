"""Small host helpers shared by the trainers (reference: miscc/utils.py keeps the same public name)."""
import os


def mkdir_p(path):
    """`mkdir -p`: create `path` with its parents; an existing directory is fine, an existing file is an error."""
    os.makedirs(path, exist_ok=True)
