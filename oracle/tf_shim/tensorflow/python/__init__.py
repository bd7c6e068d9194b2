debug = None
