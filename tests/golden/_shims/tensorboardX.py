"""Import shim for the absent `tensorboardX` package (test tree only)."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass

    def export_scalars_to_json(self, *a, **k):
        pass

    def close(self):
        pass
