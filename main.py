"""Same flow as the reference's main.py (Config -> Simulator -> run_once), on the B200 engine."""
from dvi_ekf_b200 import Config, Simulator

if __name__ == "__main__":
    config = Config("config.yaml")
    sim = Simulator(config)
    sim.run_once()
    print(f"final DOF metric: {sim.kf.calculate_dof_metric():.3E}; last update MSE: {sim.kf.update_mse:.3E}")
