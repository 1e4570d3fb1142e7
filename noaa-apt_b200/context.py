"""Context of the reference (context.rs:100-129): progress callback.  The per-step WAV export
(`--wav-steps`) is debug tracing and is not part of the fast path; `export_steps` /
`export_resample_filtered` are therefore always False here."""


class Context:
    export_steps = False
    export_resample_filtered = False

    def __init__(self, ui_callback=None):
        self._cb = ui_callback
        self.log = []

    @staticmethod
    def decode(ui_callback=None, *_unused):
        return Context(ui_callback)

    @staticmethod
    def resample(ui_callback=None, *_unused):
        return Context(ui_callback)

    def status(self, progress, description):
        self.log.append((float(progress), str(description)))
        if self._cb:
            self._cb(float(progress), str(description))
