"""`python chat/base.py` and `chat.base.generate / decode / prompt_config / main` (reference: chat/base.py) backed by lit_parrot_b200."""
from lit_parrot_b200.chat import decode, generate, main, prompt_config  # noqa: F401

if __name__ == "__main__":
    from lit_parrot_b200.cli import CLI

    CLI(main)
