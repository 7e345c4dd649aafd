class TensorBoard:
    def configure(self, *a, **k):
        pass

    def launch(self):
        return ""
