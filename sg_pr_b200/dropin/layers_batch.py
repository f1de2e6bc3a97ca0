"""Drop-in module `layers_batch`: put this directory on PYTHONPATH (or copy the reference scripts beside it) and
`from layers_batch import ...` in eval_pair.py / eval_batch.py / main_sg.py resolves to the B200-native package."""
from sg_pr_b200.layers_batch import *  # noqa: F401,F403
