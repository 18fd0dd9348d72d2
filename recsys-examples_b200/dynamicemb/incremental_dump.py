"""`dynamicemb.incremental_dump` — import path of the reference (incremental_dump.py:25-348); the functions live in `dump_load`."""
from .dump_load import get_score, incremental_dump, is_valid_score_threshold, set_score  # noqa: F401
