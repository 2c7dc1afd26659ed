import algorithm.nn_models as m

ModelRep = m.ModelSimpleRep
ModelQ = m.ModelQ
ModelPolicy = m.ModelPolicy
