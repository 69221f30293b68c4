"""Stand-in for multiprocessing_logging (self_play_parallel.py:18)."""


def install_mp_handler(*a, **k):
    pass
