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
        """The table ``fit(v=True)`` prints, in the reference's layout (``report.py:142-176``): nine columns of width ten,
        the execution time as hh:mm:s.mmm."""
        titles = ("#points", "r", "c", "min", "max", "#clusters", "%largest", "%noise")
        cells = []
        for attr in self.__slots__[:-1]:
            v = getattr(self, attr)
            cells.append(f"{v:<10.3f}" if isinstance(v, float) else f"{v!r:<10}" if v is None else f"{v:<10}")
        if self.execution_time is None:
            clock = None
        else:
            minutes, seconds = divmod(self.execution_time, 60)
            clock = f"{int(minutes // 60):0>2}:{int(minutes % 60):0>2}:{seconds:.3f}"
        rule = "-" * 95
        return f"{rule}\n{''.join(f'{t:<10}' for t in titles)}time     \n{''.join(cells)}{clock}\n{rule}\n"

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
        """One row per record; the integer columns are nullable (``member_cutoff`` / ``max_clusters`` may be None), as in
        the reference's typed frame (``report.py:47-88``)."""
        import pandas as pd

        frame = pd.DataFrame([r.to_dict() for r in self._list], columns=Record.__slots__)
        integer = ("n_points", "cnn_cutoff", "member_cutoff", "max_clusters", "n_clusters")
        return frame.astype({c: pd.Int64Dtype() if c in integer else "float64" for c in Record.__slots__})
