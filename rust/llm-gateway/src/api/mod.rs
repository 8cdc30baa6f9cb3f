pub mod rest;
