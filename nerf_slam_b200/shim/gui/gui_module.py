"""import target only: the Open3D GUI (--gui) is outside the hot-path scope"""


class GuiModule:
    def __init__(self, *a, **k):
        raise NotImplementedError("the Open3D GUI (--gui) is outside the hot-path scope; run without --gui")
