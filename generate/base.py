"""`generate.base.generate` / `main` (reference: generate/base.py:92-159, 162-257) backed by lit_parrot_b200;
`python generate/base.py --checkpoint_dir ... "prompt"` is the reference's CLI."""
from lit_parrot_b200.cli import main  # noqa: F401
from lit_parrot_b200.generate import generate  # noqa: F401

if __name__ == "__main__":
    from lit_parrot_b200.cli import CLI

    CLI(main)
