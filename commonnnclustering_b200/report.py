"""Per-fit records -- mirror of the reference's ``report.py`` (``Record`` :91-191, ``Summary`` :13-88)."""
from collections.abc import MutableSequence


class Record:
    __slots__ = ["n_points", "radius_cutoff", "cnn_cutoff", "member_cutoff", "max_clusters",
                 "n_clusters", "ratio_largest", "ratio_noise", "execution_time"]

    def __init__(self, n_points=None, radius_cutoff=None, cnn_cutoff=None, member_cutoff=None,
                 max_clusters=None, n_clusters=None, ratio_largest=None, ratio_noise=None,
                 execution_time=None):
        self.n_points = n_points
        self.radius_cutoff = radius_cutoff
        self.cnn_cutoff = cnn_cutoff
        self.member_cutoff = member_cutoff
        self.max_clusters = max_clusters
        self.n_clusters = n_clusters
        self.ratio_largest = ratio_largest
        self.ratio_noise = ratio_noise
        self.execution_time = execution_time

    def __repr__(self):
        attrs = ", ".join(f"{a}={getattr(self, a)!r}" for a in self.__slots__)
        return f"{type(self).__name__}({attrs})"

    def __str__(self):
        head = "-" * 95 + "\n" + "".join(f"{h:<10}" for h in (
            "#points", "r", "c", "min", "max", "#clusters", "%largest", "%noise", "time")) + "\n"
        vals = []
        for a in self.__slots__:
            v = getattr(self, a)
            if v is None:
                vals.append(f"{'None':<10}")
            elif isinstance(v, float):
                vals.append(f"{v:<10.3f}")
            else:
                vals.append(f"{v:<10}")
        return head + "".join(vals) + "\n" + "-" * 95 + "\n"

    def to_dict(self):
        return {a: getattr(self, a) for a in self.__slots__}


class Summary(MutableSequence):
    """List of ``Record`` objects."""

    def __init__(self, iterable=None):
        self._list = []
        if iterable is not None:
            for i in iterable:
                self.append(i)

    def __getitem__(self, key):
        return self._list[key]

    def __setitem__(self, key, item):
        if not isinstance(item, Record):
            raise TypeError("Summary can only contain records of type `Record`")
        self._list[key] = item

    def __delitem__(self, key):
        del self._list[key]

    def __len__(self):
        return len(self._list)

    def insert(self, index, item):
        if not isinstance(item, Record):
            raise TypeError("Summary can only contain records of type `Record`")
        self._list.insert(index, item)

    def __str__(self):
        return str(self._list)

    def to_DataFrame(self):
        import pandas as pd

        return pd.DataFrame([r.to_dict() for r in self._list], columns=Record.__slots__)
