"""The reference's gym.Env / VecEnv surface over the CUDA step kernel (TradingEnvironment, ModelDynamics, adapters)."""
