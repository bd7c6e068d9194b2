def get_current_tower_context():
    raise NotImplementedError('tf_shim: training-only')
