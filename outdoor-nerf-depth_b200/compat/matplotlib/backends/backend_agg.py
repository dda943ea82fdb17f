from matplotlib import _Missing
FigureCanvasAgg = _Missing("matplotlib.backends.backend_agg.FigureCanvasAgg")
