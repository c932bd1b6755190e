"""create_model(name, hparams) — same contract as tacotron/models/__init__.py:4-8 of the reference."""
from .tacotron import Tacotron


def create_model(name, hparams):
    if name == "Tacotron":
        return Tacotron(hparams)
    raise Exception("Unknown model: " + name)
