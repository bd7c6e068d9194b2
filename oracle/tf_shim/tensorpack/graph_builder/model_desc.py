class ModelDesc(object):
    pass


class InputDesc(object):
    def __init__(self, *a, **k):
        pass
