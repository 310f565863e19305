"""Empty stand-in: trainer.py:21 imports tensorflow but the step never uses it."""
