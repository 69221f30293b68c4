"""Stand-in for anytree.NodeMixin (test infrastructure only; never imported by the engine).

Only the four members that /root/reference/games/algos/mcts.py:11,21 touches are provided:
``parent`` (settable), ``children`` (ordered tuple, settable), ``is_leaf`` and the
``_post_detach_children`` hook.  anytree==2.8.0 is not installable here (no network).
"""


class NodeMixin:
    @property
    def parent(self):
        return self.__dict__.get("_nm_parent")

    @parent.setter
    def parent(self, value):
        old = self.__dict__.get("_nm_parent")
        if old is not None:
            old.__dict__["_nm_children"].remove(self)
        self.__dict__["_nm_parent"] = value
        if value is not None:
            value.__dict__.setdefault("_nm_children", []).append(self)

    @property
    def children(self):
        return tuple(self.__dict__.get("_nm_children", ()))

    @children.setter
    def children(self, nodes):
        old = self.__dict__.get("_nm_children", [])
        for c in list(old):
            c.__dict__["_nm_parent"] = None
        self._post_detach_children(tuple(old))
        self.__dict__["_nm_children"] = []
        for c in nodes:
            c.parent = self

    @property
    def is_leaf(self):
        return len(self.__dict__.get("_nm_children", ())) == 0

    def _post_detach_children(self, children):
        pass
