class BaseCallback:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k


class EvalCallback(BaseCallback):
    def __init__(self, eval_env, *a, **k):
        super().__init__(*a, **k)
        self.eval_env = eval_env
