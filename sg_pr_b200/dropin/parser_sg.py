"""Drop-in module `parser_sg`: put this directory on PYTHONPATH (or copy the reference scripts beside it) and
`from parser_sg import ...` in eval_pair.py / eval_batch.py / main_sg.py resolves to the B200-native package."""
from sg_pr_b200.parser_sg import *  # noqa: F401,F403
