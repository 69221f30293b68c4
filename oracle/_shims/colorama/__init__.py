"""Stand-in for colorama (render() only; connect4env.py:3-5)."""


def init(*a, **k):
    pass


class _Blank:
    def __getattr__(self, name):
        return ""


Fore = _Blank()
Style = _Blank()
