pub mod dto;
pub mod handlers;
pub mod routes;
