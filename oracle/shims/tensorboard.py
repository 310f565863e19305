"""Stand-in for the pre-2018 `tensorboard` package trainer.py:19-20 imports."""
from tensorboardX import summary, FileWriter  # noqa: F401
