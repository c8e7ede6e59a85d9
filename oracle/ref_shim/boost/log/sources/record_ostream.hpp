// Shim: intentionally empty (see severity_logger.hpp in this directory).
#pragma once
