pub mod service;
