"""Stand-in for tensorboardX (absent here): the two names the reference trainer imports."""


class _Summary:
    @staticmethod
    def scalar(name, value):
        return (name, float(value))


summary = _Summary()


class FileWriter:
    def __init__(self, logdir=None):
        self.events = []

    def add_summary(self, s, step=None):
        self.events.append((step, s))

    def flush(self):
        pass

    def close(self):
        pass
