"""wandb stand-in -- TEST INFRASTRUCTURE ONLY (experiment logging is outside the hot path)."""


class _Run:
    id = "offline"

    def finish(self):
        pass


def init(*a, **k):
    return _Run()


def log(*a, **k):
    pass
