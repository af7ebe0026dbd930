"""`python generate/full.py` (reference: generate/full.py) backed by lit_parrot_b200."""
from lit_parrot_b200.cli_finetuned import generate_prompt  # noqa: F401
from lit_parrot_b200.cli_finetuned import main_full as main  # noqa: F401
from lit_parrot_b200.generate import generate  # noqa: F401

if __name__ == "__main__":
    from lit_parrot_b200.cli import CLI

    CLI(main)
